// b2p_dense.cu -- the dense LCP of dWorldStep (fast_step == false) for sm_100a.
//
// The reference selects ODE's "big matrix" stepper with fast_step == false (src/WorldPhysics.cpp:367-370: dWorldStep
// instead of dWorldQuickStep; the WorldPhysics constructor's default, src/WorldPhysics.cpp:82).  ODE's step.cpp builds
// A = J invM J^T + diag(cfm / h), rhs = c/h - J (v/h + invM fe) per island and solves the box-LCP exactly with
// dSolveLCP (lcp.cpp): Dantzig's principal pivoting as described by Baraff (SIGGRAPH 94).  oracle/orc_lcp.c restates
// it on the CPU; this kernel is the same algorithm, step by step, for one world per thread:
//   * the rows come from the assemble kernel exactly as the SOR kernels get them (b2p_assemble.cu: scaled rows
//     Jh = sqrt(Ad) J M^-1/2, rhs', cfm', one bound word, in static slots; `ord0` lists the used slots with the
//     friction-dependent rows last).  The scaling is a positive diagonal change of variables (lam' = lam / sqrt(Ad),
//     w' = sqrt(Ad) w), which keeps every sign test and every step ratio of the pivoting method, so the scaled
//     problem A' = Jh Jh^T + diag(cfm') is solved directly and its multipliers / fch = Jh^T lam' go to the
//     integrate kernel unchanged;
//   * unbounded rows enter the clamped set C first and are solved at once; when the first friction-dependent row is
//     reached all of them get their bounds +-|mu' lam'_n| from the friction-less solution found so far (ODE:
//     hit_first_friction_index); every further row is driven until it is in C or at a bound, rows changing set as
//     the step length dictates (cases 1..6 of lcp.cpp);
//   * ODE updates an LDL^T factor of A(C,C) row by row; so does the warp kernel (a row is appended when an index
//     enters C, the rows behind the position of an index that leaves are rebuilt); the thread-per-world kernel and the
//     oracle keep a Cholesky factor -- the same solves up to rounding;
//   * islands: A is block diagonal over a world's islands, so one problem per world has the same solution as one
//     per island, whichever row order (world block / ODE islands) the assemble kernel used.
// Arithmetic in FP64 (ODE's dWorldStep is used with dDOUBLE; a Cholesky factor of a constraint matrix whose only
// regularisation is cfm ~ 1e-5 .. 1e-10 is not an FP32 computation), rows and results FP32 like the rest of the step.
// Two kernels with the same pivoting sequence: b2p_dense_warp_kernel (one warp per world, the whole problem in shared
// memory; the one that runs for scenes up to ~160 row slots) and b2p_dense_lcp_kernel (one thread per world, per-world
// scratch in global memory with the world index fastest; the fallback and the cross-check).  This is the exact,
// O(m^3) stepper -- tens of times slower than the SOR kernels by construction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>

#include "b2p_dev.h"
#include "b2p_math.cuh"

namespace {

constexpr int DBD = 64; // threads (= worlds) per CTA

#define DV(arr, i) (arr)[(size_t)(i) * Wp + w]

__global__ void __launch_bounds__(DBD) b2p_dense_lcp_kernel(const StepParams P, const DenseScratch D)
{
    const DevTemplate *__restrict__ T = P.T;
    const int w = blockIdx.x * DBD + threadIdx.x;
    if (w >= P.W) return;
    const size_t Wp = (size_t)P.Wp;
    const int R = T->maxrows, nb = T->nb;
    const int n = P.nrows[w] & 0xffff;
    double *const A = D.A, *const L = D.L;
    double *const x = D.vec, *const wv = x + (size_t)R * Wp, *const lo = wv + (size_t)R * Wp,
                 *const hi = lo + (size_t)R * Wp, *const dx = hi + (size_t)R * Wp, *const dw = dx + (size_t)R * Wp,
                 *const bv = dw + (size_t)R * Wp, *const rp = bv + (size_t)R * Wp, *const sp = rp + (size_t)R * Wp,
                 *const fc = sp + (size_t)R * Wp; // fc: [(nb + 1) * 6]
    uint8_t *const Cl = D.sets, *const set = Cl + (size_t)R * Wp, *const Cold = set + (size_t)R * Wp,
                  *const fidx = Cold + (size_t)R * Wp; // fidx: position of the normal row, 0xff: none

    // ---- the rows in processing order (positions p = 0 .. n-1 of ord0): A, b, bounds
    const char *const rows_w = (const char *)ROW_UNIT(P.rows, 0, 0, Wp, w);
    const size_t slot_stride = Wp * 64;
    struct RowD {
        float j[12];
        float rhs, cfm, bound;
        uint32_t meta;
    };
    auto load_row = [&](int p) -> RowD {
        const int slot = P.ord0[(size_t)p * Wp + w];
        const char *const pu = rows_w + (size_t)slot * slot_stride;
        const F8 u0 = ldg256(pu), u1 = ldg256(pu + (slot_stride >> 1));
        RowD q;
        q.j[0] = u0.lo.x, q.j[1] = u0.lo.y, q.j[2] = u0.lo.z, q.j[3] = u0.lo.w;
        q.j[4] = u0.hi.x, q.j[5] = u0.hi.y, q.j[6] = u0.hi.z, q.j[7] = u0.hi.w;
        q.j[8] = u1.lo.x, q.j[9] = u1.lo.y, q.j[10] = u1.lo.z, q.j[11] = u1.lo.w;
        q.rhs = u1.hi.x, q.cfm = u1.hi.y, q.bound = u1.hi.z;
        q.meta = (uint32_t)__float_as_int(u1.hi.w);
        return q;
    };
    for (int k = 0; k < (nb + 1) * 6; k++) DV(fc, k) = 0.0;
    int nub = 0;
    for (int p = 0; p < n; p++) {
        const RowD r = load_row(p);
        const int a0 = (r.meta >> 16) & 0x3f, a1 = (r.meta >> 24) & 0x3f;
        for (int q = 0; q <= p; q++) {
            const RowD s = load_row(q);
            const int b0 = (s.meta >> 16) & 0x3f, b1 = (s.meta >> 24) & 0x3f;
            double acc = 0.0;
            if (a0 == b0) for (int k = 0; k < 6; k++) acc += (double)r.j[k] * (double)s.j[k];
            if (a0 == b1) for (int k = 0; k < 6; k++) acc += (double)r.j[k] * (double)s.j[6 + k];
            if (a1 == b0) for (int k = 0; k < 6; k++) acc += (double)r.j[6 + k] * (double)s.j[k];
            if (a1 == b1) for (int k = 0; k < 6; k++) acc += (double)r.j[6 + k] * (double)s.j[6 + k];
            if (q == p) acc += (double)r.cfm;
            DV(A, p * R + q) = acc;
            DV(A, q * R + p) = acc;
        }
        DV(bv, p) = (double)r.rhs;
        const int fslot = (r.meta >> 8) & 0xff;
        const double bnd = (double)r.bound;
        double l = -bnd, h = bnd;
        if (r.meta & 0x40000000u) l = 0.0;
        if (r.meta & 0x80000000u) h = 0.0;
        DV(lo, p) = l;
        DV(hi, p) = h;
        DV(x, p) = 0.0;
        DV(wv, p) = 0.0;
        DV(set, p) = 0;
        DV(fidx, p) = (uint8_t)(fslot != R + 1 ? fslot : 0xff); // slot of the normal row for now
    }
    // friction-dependent rows: the POSITION of the normal row (a world-block or an island order may put it anywhere)
    for (int p = 0; p < n; p++) {
        const int fslot = DV(fidx, p);
        if (fslot == 0xff) continue;
        uint8_t fp = 0xff;
        for (int q = 0; q < n; q++)
            if (P.ord0[(size_t)q * Wp + w] == fslot) fp = (uint8_t)q;
        DV(fidx, p) = fp;
    }

    // ---- Cholesky factor of A(C,C) in position space
    int nC = 0;
    bool failed = false;
    auto chol_append = [&](int idx) {
        const int p = nC;
        double d = DV(A, idx * R + idx);
        for (int q = 0; q < p; q++) {
            double s = DV(A, idx * R + DV(Cl, q));
            for (int r = 0; r < q; r++) s -= DV(L, q * R + r) * DV(L, p * R + r);
            const double v = s / DV(L, q * R + q);
            DV(L, p * R + q) = v;
            d -= v * v;
        }
        if (!(d > 0.0)) {
            failed = true;
            d = 1e-300;
        }
        DV(L, p * R + p) = sqrt(d);
        DV(Cl, p) = (uint8_t)idx;
        nC = p + 1;
    };
    auto chol_remove = [&](int idx) {
        const int old = nC;
        for (int q = 0; q < old; q++) DV(Cold, q) = DV(Cl, q);
        nC = 0;
        for (int q = 0; q < old; q++) {
            const int c = DV(Cold, q);
            if (c != idx) chol_append(c);
        }
    };
    // sp(C) = A(C,C)^-1 rp(C), by position
    auto chol_solve = [&]() {
        for (int q = 0; q < nC; q++) {
            double s = DV(rp, q);
            for (int r = 0; r < q; r++) s -= DV(L, q * R + r) * DV(sp, r);
            DV(sp, q) = s / DV(L, q * R + q);
        }
        for (int q = nC - 1; q >= 0; q--) {
            double s = DV(sp, q);
            for (int r = q + 1; r < nC; r++) s -= DV(L, r * R + q) * DV(sp, r);
            DV(sp, q) = s / DV(L, q * R + q);
        }
    };

    // ---- processing order (ODE's dLCP constructor): unbounded rows, then the bounded ones, then the rows whose
    // bounds follow a normal row's multiplier
    uint8_t *const perm = fidx + (size_t)R * Wp;
    {
        int k = 0;
        for (int p = 0; p < n; p++)
            if (DV(fidx, p) == 0xff && DV(lo, p) == -INFINITY && DV(hi, p) == INFINITY) DV(perm, k++) = (uint8_t)p;
        nub = k;
        for (int p = 0; p < n; p++)
            if (DV(fidx, p) == 0xff && !(DV(lo, p) == -INFINITY && DV(hi, p) == INFINITY)) DV(perm, k++) = (uint8_t)p;
        for (int p = 0; p < n; p++)
            if (DV(fidx, p) != 0xff) DV(perm, k++) = (uint8_t)p;
    }
    for (int k = 0; k < nub; k++) {
        const int i = DV(perm, k);
        chol_append(i);
        DV(set, i) = 1;
    }
    if (nub > 0) {
        for (int q = 0; q < nub; q++) DV(rp, q) = DV(bv, DV(Cl, q));
        chol_solve();
        for (int q = 0; q < nub; q++) DV(x, DV(Cl, q)) = DV(sp, q);
    }

    bool err = false, hit_first_friction_index = false;
    int guard = 0;
    const int guard_max = 50 * n + 100;
    for (int k = nub; k < n && !err; k++) {
        const int i = DV(perm, k);
        if (!hit_first_friction_index && DV(fidx, i) != 0xff) {
            for (int kk = k; kk < n; kk++) {
                const int r = DV(perm, kk);
                const double wfk = DV(x, DV(fidx, r));
                if (wfk == 0.0) {
                    DV(hi, r) = 0.0, DV(lo, r) = 0.0;
                } else {
                    const double h = fabs(DV(hi, r) * wfk);
                    DV(hi, r) = h, DV(lo, r) = -h;
                }
            }
            hit_first_friction_index = true;
        }
        {
            double s = -DV(bv, i);
            for (int j = 0; j < n; j++)
                if (DV(set, j)) s += DV(A, i * R + j) * DV(x, j);
            DV(wv, i) = s;
        }
        const double wi0 = DV(wv, i);
        if (DV(lo, i) == 0.0 && wi0 >= 0.0) {
            DV(set, i) = 2;
            continue;
        }
        if (DV(hi, i) == 0.0 && wi0 <= 0.0) {
            DV(set, i) = 3;
            continue;
        }
        if (wi0 == 0.0) {
            chol_append(i);
            DV(set, i) = 1;
            continue;
        }
        for (;;) {
            if (++guard > guard_max) {
                err = true;
                break;
            }
            const double dirf = DV(wv, i) <= 0.0 ? 1.0 : -1.0;
            for (int q = 0; q < nC; q++) DV(rp, q) = -dirf * DV(A, (int)DV(Cl, q) * R + i);
            chol_solve();
            for (int q = 0; q < nC; q++) DV(dx, DV(Cl, q)) = DV(sp, q);
            for (int j = 0; j < n; j++) {
                if (!(DV(set, j) >= 2 || j == i)) continue;
                double s = DV(A, j * R + i) * dirf;
                for (int q = 0; q < nC; q++) s += DV(A, j * R + DV(Cl, q)) * DV(sp, q);
                DV(dw, j) = s;
            }
            int cmd = 1, si = 0;
            double s = -DV(wv, i) / DV(dw, i);
            if (dirf > 0.0) {
                if (DV(hi, i) < INFINITY) {
                    const double s2 = (DV(hi, i) - DV(x, i)) * dirf;
                    if (s2 < s) s = s2, cmd = 3;
                }
            } else {
                if (DV(lo, i) > -INFINITY) {
                    const double s2 = (DV(lo, i) - DV(x, i)) * dirf;
                    if (s2 < s) s = s2, cmd = 2;
                }
            }
            for (int j = 0; j < n; j++) {
                const int st = DV(set, j);
                if (st < 2) continue;
                const double dwj = DV(dw, j);
                if ((st == 2 && dwj < 0.0) || (st == 3 && dwj > 0.0)) {
                    if (DV(lo, j) == 0.0 && DV(hi, j) == 0.0) continue;
                    const double s2 = -DV(wv, j) / dwj;
                    if (s2 < s) s = s2, cmd = 4, si = j;
                }
            }
            for (int q = nub; q < nC; q++) {
                const int j = DV(Cl, q);
                const double dxj = DV(dx, j);
                if (dxj < 0.0 && DV(lo, j) > -INFINITY) {
                    const double s2 = (DV(lo, j) - DV(x, j)) / dxj;
                    if (s2 < s) s = s2, cmd = 5, si = j;
                }
                if (dxj > 0.0 && DV(hi, j) < INFINITY) {
                    const double s2 = (DV(hi, j) - DV(x, j)) / dxj;
                    if (s2 < s) s = s2, cmd = 6, si = j;
                }
            }
            if (!(s > 0.0)) {
                // ODE: "LCP internal error, s <= 0": give up with the current solution
                for (int kk = k; kk < n; kk++) DV(x, DV(perm, kk)) = 0.0, DV(wv, DV(perm, kk)) = 0.0;
                err = true;
                break;
            }
            for (int q = 0; q < nC; q++) DV(x, DV(Cl, q)) += s * DV(dx, DV(Cl, q));
            DV(x, i) += s * dirf;
            for (int j = 0; j < n; j++)
                if (DV(set, j) >= 2) DV(wv, j) += s * DV(dw, j);
            DV(wv, i) += s * DV(dw, i);
            switch (cmd) {
            case 1: DV(wv, i) = 0.0; chol_append(i); DV(set, i) = 1; break;
            case 2: DV(x, i) = DV(lo, i); DV(set, i) = 2; break;
            case 3: DV(x, i) = DV(hi, i); DV(set, i) = 3; break;
            case 4: DV(wv, si) = 0.0; chol_append(si); DV(set, si) = 1; break;
            case 5: DV(x, si) = DV(lo, si); DV(set, si) = 2; chol_remove(si); break;
            case 6: DV(x, si) = DV(hi, si); DV(set, si) = 3; chol_remove(si); break;
            }
            if (cmd <= 3) break;
        }
    }
    if ((err || failed) && D.failures) atomicAdd(D.failures, 1);

    // ---- results where the integrate kernel expects them: multipliers by slot, fch = Jh^T lam' per body
    double resid = 0.0;
    for (int p = 0; p < n; p++) {
        const RowD r = load_row(p);
        const double xp = DV(x, p);
        const int slot = r.meta & 0xff, a0 = (r.meta >> 16) & 0x3f, a1 = (r.meta >> 24) & 0x3f;
        P.lambda[(size_t)slot * Wp + w] = (float)xp;
        for (int k = 0; k < 6; k++) {
            DV(fc, a0 * 6 + k) += xp * (double)r.j[k];
            DV(fc, a1 * 6 + k) += xp * (double)r.j[6 + k];
        }
        if (D.residual) {
            // complementarity residual of the scaled problem (debug): |w| inside the bounds, the wrong-signed part at one
            double s = -DV(bv, p);
            for (int j = 0; j < n; j++) s += DV(A, p * R + j) * DV(x, j);
            const double l = DV(lo, p), h = DV(hi, p);
            double res;
            if (l == h) res = 0.0;
            else if (xp <= l && l > -INFINITY) res = s < 0.0 ? -s : 0.0;
            else if (xp >= h && h < INFINITY) res = s > 0.0 ? s : 0.0;
            else res = fabs(s);
            resid = res > resid ? res : resid;
        }
    }
    if (D.residual) D.residual[w] = (float)resid;
    for (int b = 0; b < nb; b++)
        for (int k = 0; k < 6; k++) P.fcacc[(size_t)(b * 6 + k) * Wp + w] = (float)DV(fc, b * 6 + k);
}


// ---------------------------------------------------------------------------------------------------------------
// The same algorithm with ONE WARP per world and the whole problem in shared memory: the matrix and its factor packed
// as lower triangles, the vectors, the rows' Jacobians and the index sets -- 26 KB per world at the rover's 48 row
// slots, so eight worlds are in flight per SM, each advanced by 32 lanes: position vectors live in registers
// (element r of a vector in lane r mod 32), the triangular solves run column by column with one shuffle broadcast
// per column, the matrix-vector products and the step-length search are split over the lanes (lexicographic
// minimum over (step, scan position), so ties resolve as in the serial scan).  Used whenever a world's arrays fit
// the shared memory of an SM (up to about 160 row slots); the kernel above is the fallback and the cross-check.
constexpr unsigned FULLW = 0xffffffffu;

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; } // i >= j
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULLW, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLW, v, o));
    return v;
}

template <int K> // position-vector elements per lane: ceil(maxrows / 32)
__global__ void __launch_bounds__(32) b2p_dense_warp_kernel(const StepParams P, const DenseScratch D)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const DevTemplate *__restrict__ T = P.T;
    const int lane = threadIdx.x;
    const size_t Wp = (size_t)P.Wp;
    const int R = T->maxrows, nb = T->nb, NT = R * (R + 1) / 2;
    double *const As = (double *)smraw, *const Ls = As + NT, *const invd = Ls + NT, *const x = invd + R,
                 *const wv = x + R, *const lo = wv + R, *const hi = lo + R, *const dw = hi + R, *const bv = dw + R,
                 *const sp = bv + R, *const cf = sp + R; // cf: cfm' of the rows while A is built
    float *const rowJ = (float *)(cf + R); // [R][12]
    uint8_t *const ordS = (uint8_t *)(rowJ + (size_t)R * 12), *const set = ordS + R, *const Cl = set + R,
                  *const fidx = Cl + R, *const perm = fidx + R, *const Cold = perm + R, *const bod = Cold + R; // bod [2R]
    auto Asym = [&](int i, int j) -> double { return As[i >= j ? tri(i, j) : tri(j, i)]; };

    for (int w = blockIdx.x; w < P.W; w += gridDim.x) {
        __syncwarp();
        const int n = P.nrows[w] & 0xffff;
        const char *const rows_w = (const char *)ROW_UNIT(P.rows, 0, 0, Wp, w);
        const size_t slot_stride = Wp * 64;
        // ---- rows into shared memory (position p = place in ord0's order)
        for (int p = lane; p < n; p += 32) {
            const int slot = P.ord0[(size_t)p * Wp + w];
            const char *const pu = rows_w + (size_t)slot * slot_stride;
            const F8 u0 = ldg256(pu), u1 = ldg256(pu + (slot_stride >> 1));
            float *j = rowJ + p * 12;
            j[0] = u0.lo.x, j[1] = u0.lo.y, j[2] = u0.lo.z, j[3] = u0.lo.w, j[4] = u0.hi.x, j[5] = u0.hi.y;
            j[6] = u0.hi.z, j[7] = u0.hi.w, j[8] = u1.lo.x, j[9] = u1.lo.y, j[10] = u1.lo.z, j[11] = u1.lo.w;
            const uint32_t meta = (uint32_t)__float_as_int(u1.hi.w);
            bv[p] = (double)u1.hi.x;
            cf[p] = (double)u1.hi.y;
            const double bnd = (double)u1.hi.z;
            lo[p] = (meta & 0x40000000u) ? 0.0 : -bnd;
            hi[p] = (meta & 0x80000000u) ? 0.0 : bnd;
            x[p] = 0.0, wv[p] = 0.0;
            set[p] = 0;
            ordS[p] = (uint8_t)slot;
            const int fslot = (meta >> 8) & 0xff;
            fidx[p] = (uint8_t)(fslot != R + 1 ? fslot : 0xff);
            bod[2 * p] = (uint8_t)((meta >> 16) & 0x3f), bod[2 * p + 1] = (uint8_t)((meta >> 24) & 0x3f);
        }
        __syncwarp();
        // ---- A' = Jh Jh^T + diag(cfm'), lower triangle, one entry per lane and pass
        for (int e = lane; e < n * (n + 1) / 2; e += 32) {
            int p = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while (tri(p + 1, 0) <= e) p++;
            while (tri(p, 0) > e) p--;
            const int q = e - tri(p, 0);
            const float *r = rowJ + p * 12, *t = rowJ + q * 12;
            const int a0 = bod[2 * p], a1 = bod[2 * p + 1], b0 = bod[2 * q], b1 = bod[2 * q + 1];
            double acc = 0.0;
            if (a0 == b0) for (int k = 0; k < 6; k++) acc += (double)r[k] * (double)t[k];
            if (a0 == b1) for (int k = 0; k < 6; k++) acc += (double)r[k] * (double)t[6 + k];
            if (a1 == b0) for (int k = 0; k < 6; k++) acc += (double)r[6 + k] * (double)t[k];
            if (a1 == b1) for (int k = 0; k < 6; k++) acc += (double)r[6 + k] * (double)t[6 + k];
            if (q == p) acc += cf[p];
            As[e] = acc;
        }
        // friction-dependent rows: the position of the normal row
        for (int p = lane; p < n; p += 32) {
            const int fslot = fidx[p];
            if (fslot == 0xff) continue;
            uint8_t fp = 0xff;
            for (int q = 0; q < n; q++)
                if (ordS[q] == fslot) fp = (uint8_t)q;
            fidx[p] = fp;
        }
        __syncwarp();
        // ---- processing order: unbounded rows, the bounded ones, the friction-dependent ones (stable)
        int nub = 0;
        {
            int cnt = 0;
            for (int pass = 0; pass < 3; pass++) {
                for (int base = 0; base < n; base += 32) {
                    const int p = base + lane;
                    bool pr = false;
                    if (p < n) {
                        const bool dep = fidx[p] != 0xff, unb = lo[p] == -INFINITY && hi[p] == INFINITY;
                        pr = (dep ? 2 : (unb ? 0 : 1)) == pass;
                    }
                    const unsigned m = __ballot_sync(FULLW, pr);
                    if (pr) perm[cnt + __popc(m & ((1u << lane) - 1u))] = (uint8_t)p;
                    cnt += __popc(m);
                }
                if (pass == 0) nub = cnt;
            }
        }
        __syncwarp();

        int nC = 0;
        bool failed = false;
        // The factor is L D L^T (unit lower L, like ODE's): a column of a triangular solve is then one shuffle
        // broadcast and one DFMA per element, with no multiplication by the diagonal on the dependent chain.
        // Forward / backward substitution with the factor's first p rows on a position vector held one element per
        // lane and 32 positions (in place); `invd` holds 1 / D.
        int roff[K]; // offset of this lane's rows in the packed triangle
#pragma unroll
        for (int k = 0; k < K; k++) roff[k] = tri(lane + 32 * k, 0);
        auto fwd = [&](double(&y)[K], int p) {
            for (int q = 0; q < p; q++) {
                const int kq = q >> 5, owner = q & 31;
                double v = y[0];
#pragma unroll
                for (int k = 1; k < K; k++)
                    if (k == kq) v = y[k];
                v = __shfl_sync(FULLW, v, owner);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int r = lane + 32 * k;
                    if (r > q && r < p) y[k] -= Ls[roff[k] + q] * v;
                }
            }
        };
        auto bwd = [&](double(&y)[K], int p) {
            for (int q = p - 1; q >= 0; q--) {
                const int kq = q >> 5, owner = q & 31, qoff = tri(q, 0);
                double v = y[0];
#pragma unroll
                for (int k = 1; k < K; k++)
                    if (k == kq) v = y[k];
                v = __shfl_sync(FULLW, v, owner);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int r = lane + 32 * k;
                    if (r < q) y[k] -= Ls[qoff + r] * v;
                }
            }
        };
        // new last row of the factor from u = L^-1 A(C, idx) (by position): l = u / D, d = A(idx,idx) - u . l
        auto factor_append_row = [&](int idx, const double(&u)[K]) {
            const int p = nC;
            double ss = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int r = lane + 32 * k;
                if (r < p) {
                    const double l = u[k] * invd[r];
                    ss += u[k] * l;
                    Ls[tri(p, r)] = l;
                }
            }
            double d = As[tri(idx, idx)] - warp_sum(ss);
            if (!(d > 0.0)) {
                failed = true;
                d = 1e-300;
            }
            if (lane == 0) {
                invd[p] = 1.0 / d;
                Cl[p] = (uint8_t)idx;
            }
            nC = p + 1;
            __syncwarp();
        };
        auto factor_append = [&](int idx) {
            const int p = nC;
            double y[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int r = lane + 32 * k;
                y[k] = r < p ? Asym(idx, Cl[r]) : 0.0;
            }
            fwd(y, p);
            factor_append_row(idx, y);
        };
        // an index leaves C: the factor's rows in front of its position stay, the ones behind are rebuilt
        auto factor_remove = [&](int idx) {
            const int old = nC;
            int pos = old;
            for (int base = 0; base < old; base += 32) {
                const int q = base + lane;
                const unsigned m = __ballot_sync(FULLW, q < old && Cl[q] == idx);
                if (m) pos = base + __ffs((int)m) - 1;
            }
            for (int q = pos + 1 + lane; q < old; q += 32) Cold[q] = Cl[q];
            __syncwarp();
            nC = pos;
            for (int q = pos + 1; q < old; q++) factor_append(Cold[q]);
        };
        // sp(C) = A(C,C)^-1 y, y by position (one element per lane and 32 positions)
        double zfwd[K]; // L^-1 rhs of the last solve: if the driving index enters C next, that is its row of the factor
        auto solve_to_sp = [&](double(&y)[K]) {
            fwd(y, nC);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int r = lane + 32 * k;
                zfwd[k] = y[k];
                if (r < nC) y[k] *= invd[r];
            }
            bwd(y, nC);
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int r = lane + 32 * k;
                if (r < nC) sp[r] = y[k];
            }
            __syncwarp();
        };

        for (int k = 0; k < nub; k++) {
            const int i = perm[k];
            factor_append(i);
            if (lane == 0) set[i] = 1;
        }
        __syncwarp();
        if (nub > 0) {
            double y[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int r = lane + 32 * k;
                y[k] = r < nC ? bv[Cl[r]] : 0.0;
            }
            solve_to_sp(y);
            for (int q = lane; q < nC; q += 32) x[Cl[q]] = sp[q];
            __syncwarp();
        }

        bool err = false, hit_first_friction_index = false;
        int guard = 0;
        const int guard_max = 50 * n + 100;
        for (int k = nub; k < n && !err; k++) {
            const int i = perm[k];
            if (!hit_first_friction_index && fidx[i] != 0xff) {
                for (int kk = k + lane; kk < n; kk += 32) {
                    const int r = perm[kk];
                    const double wfk = x[fidx[r]];
                    if (wfk == 0.0) {
                        hi[r] = 0.0, lo[r] = 0.0;
                    } else {
                        const double h = fabs(hi[r] * wfk);
                        hi[r] = h, lo[r] = -h;
                    }
                }
                hit_first_friction_index = true;
                __syncwarp();
            }
            {
                double s = 0.0;
                for (int j = lane; j < n; j += 32)
                    if (set[j]) s += Asym(i, j) * x[j];
                s = warp_sum(s) - bv[i];
                if (lane == 0) wv[i] = s;
                __syncwarp();
            }
            const double wi0 = wv[i];
            if (lo[i] == 0.0 && wi0 >= 0.0) {
                if (lane == 0) set[i] = 2;
                __syncwarp();
                continue;
            }
            if (hi[i] == 0.0 && wi0 <= 0.0) {
                if (lane == 0) set[i] = 3;
                __syncwarp();
                continue;
            }
            if (wi0 == 0.0) {
                factor_append(i);
                if (lane == 0) set[i] = 1;
                __syncwarp();
                continue;
            }
            for (;;) {
                if (++guard > guard_max) {
                    err = true;
                    break;
                }
                const double dirf = wv[i] <= 0.0 ? 1.0 : -1.0;
                {
                    double y[K];
#pragma unroll
                    for (int kk = 0; kk < K; kk++) {
                        const int r = lane + 32 * kk;
                        y[kk] = r < nC ? -dirf * Asym(Cl[r], i) : 0.0;
                    }
                    solve_to_sp(y); // sp = dx(C) by position
                }
                for (int j = lane; j < n; j += 32) {
                    if (!(set[j] >= 2 || j == i)) continue;
                    double s = Asym(j, i) * dirf;
                    for (int q = 0; q < nC; q++) s += Asym(j, Cl[q]) * sp[q];
                    dw[j] = s;
                }
                __syncwarp();
                // the largest step: lexicographic minimum over (step, place in the serial scan)
                double s = -wv[i] / dw[i];
                int ord = 0, pay = 1 << 8;
                auto cand = [&](double s2, int o, int c, int sidx) {
                    if (s2 < s || (s2 == s && o < ord)) s = s2, ord = o, pay = (c << 8) | sidx;
                };
                if (dirf > 0.0) {
                    if (hi[i] < INFINITY) cand((hi[i] - x[i]) * dirf, 1, 3, 0);
                } else {
                    if (lo[i] > -INFINITY) cand((lo[i] - x[i]) * dirf, 1, 2, 0);
                }
                for (int j = lane; j < n; j += 32) {
                    const int st = set[j];
                    if (st < 2) continue;
                    const double dwj = dw[j];
                    if ((st == 2 && dwj < 0.0) || (st == 3 && dwj > 0.0)) {
                        if (lo[j] == 0.0 && hi[j] == 0.0) continue;
                        cand(-wv[j] / dwj, 2 + j, 4, j);
                    }
                }
                for (int q = nub + lane; q < nC; q += 32) {
                    const int j = Cl[q];
                    const double dxj = sp[q];
                    if (dxj < 0.0 && lo[j] > -INFINITY) cand((lo[j] - x[j]) / dxj, 2 + n + 2 * q, 5, j);
                    if (dxj > 0.0 && hi[j] < INFINITY) cand((hi[j] - x[j]) / dxj, 3 + n + 2 * q, 6, j);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const double s2 = __shfl_xor_sync(FULLW, s, o);
                    const int o2 = __shfl_xor_sync(FULLW, ord, o), p2 = __shfl_xor_sync(FULLW, pay, o);
                    if (s2 < s || (s2 == s && o2 < ord)) s = s2, ord = o2, pay = p2;
                }
                const int cmd = pay >> 8, si = pay & 0xff;
                if (!(s > 0.0)) {
                    for (int kk = k + lane; kk < n; kk += 32) x[perm[kk]] = 0.0, wv[perm[kk]] = 0.0;
                    __syncwarp();
                    err = true;
                    break;
                }
                for (int q = lane; q < nC; q += 32) x[Cl[q]] += s * sp[q];
                for (int j = lane; j < n; j += 32)
                    if (set[j] >= 2) wv[j] += s * dw[j];
                __syncwarp();
                if (lane == 0) {
                    x[i] += s * dirf;
                    wv[i] += s * dw[i];
                    switch (cmd) {
                    case 1: wv[i] = 0.0; set[i] = 1; break;
                    case 2: x[i] = lo[i]; set[i] = 2; break;
                    case 3: x[i] = hi[i]; set[i] = 3; break;
                    case 4: wv[si] = 0.0; set[si] = 1; break;
                    case 5: x[si] = lo[si]; set[si] = 2; break;
                    case 6: x[si] = hi[si]; set[si] = 3; break;
                    }
                }
                __syncwarp();
                if (cmd == 1) {
                    // the solve above ran on rhs = -dir A(C,i) with the C that i now joins: its forward half is the
                    // new row (ODE keeps `ell` from solve1 for the same reason)
                    double u[K];
#pragma unroll
                    for (int kk = 0; kk < K; kk++) u[kk] = -dirf * zfwd[kk];
                    factor_append_row(i, u);
                } else if (cmd == 4) factor_append(si);
                else if (cmd >= 5) factor_remove(si);
                if (cmd <= 3) break;
            }
        }
        if ((err || failed) && D.failures && lane == 0) atomicAdd(D.failures, 1);

        // ---- results: multipliers by slot, fch = Jh^T lam' per body (each component summed by one lane: the same
        // order in every run)
        for (int p = lane; p < n; p += 32) P.lambda[(size_t)ordS[p] * Wp + w] = (float)x[p];
        for (int c = lane; c < nb * 6; c += 32) {
            const int b = c / 6, kk = c - 6 * b;
            double acc = 0.0;
            for (int p = 0; p < n; p++) {
                if (bod[2 * p] == b) acc += x[p] * (double)rowJ[p * 12 + kk];
                if (bod[2 * p + 1] == b) acc += x[p] * (double)rowJ[p * 12 + 6 + kk];
            }
            P.fcacc[(size_t)c * Wp + w] = (float)acc;
        }
        if (D.residual) {
            double resid = 0.0;
            for (int p = lane; p < n; p += 32) {
                double s = -bv[p];
                for (int j = 0; j < n; j++) s += Asym(p, j) * x[j];
                const double l = lo[p], h = hi[p], xp = x[p];
                double res;
                if (l == h) res = 0.0;
                else if (xp <= l && l > -INFINITY) res = s < 0.0 ? -s : 0.0;
                else if (xp >= h && h < INFINITY) res = s > 0.0 ? s : 0.0;
                else res = fabs(s);
                resid = res > resid ? res : resid;
            }
            resid = warp_max(resid);
            if (lane == 0) D.residual[w] = (float)resid;
        }
    }
}

// shared memory of the warp kernel per world (= per CTA of one warp)
size_t dense_warp_smem(int nb, int maxrows)
{
    const size_t R = (size_t)maxrows, NT = R * (R + 1) / 2;
    size_t s = sizeof(double) * (2 * NT + 10 * R) + sizeof(float) * 12 * R + 8 * R;
    (void)nb;
    return (s + 15) & ~(size_t)15;
}

} // namespace

// per-world scratch of the dense stepper, in elements: doubles of A / of the factor / of the vectors, bytes of the sets
extern "C" void b2p_dense_scratch_sizes(int nb, int maxrows, size_t *mat_doubles, size_t *vec_doubles, size_t *set_bytes)
{
    *mat_doubles = (size_t)maxrows * maxrows;
    *vec_doubles = (size_t)maxrows * 9 + (size_t)(nb + 1) * 6;
    *set_bytes = (size_t)maxrows * 5;
}

// does the warp kernel (all of a world in shared memory) run for this scene?  Otherwise the thread-per-world kernel,
// which needs the per-world global scratch of b2p_dense_scratch_sizes
extern "C" int b2p_dense_fits_warp_kernel(int nb, int maxrows)
{
    const int K = (maxrows + 31) / 32;
    return dense_warp_smem(nb, maxrows) + 1024 <= 227 * 1024 && K >= 1 && K <= 5;
}

// `force_thread_kernel`: the one-world-per-thread kernel even where the warp kernel fits (tests compare the two)
extern "C" cudaError_t b2p_launch_dense(const StepParams *P, const DenseScratch *D, int nb, int maxrows, int sm_count,
                                        int force_thread_kernel, cudaStream_t stream)
{
    const size_t smem = dense_warp_smem(nb, maxrows);
    const int K = (maxrows + 31) / 32;
    if (!force_thread_kernel && b2p_dense_fits_warp_kernel(nb, maxrows)) {
        void (*kern)(const StepParams, const DenseScratch) = nullptr;
        switch (K) {
        case 1: kern = b2p_dense_warp_kernel<1>; break;
        case 2: kern = b2p_dense_warp_kernel<2>; break;
        case 3: kern = b2p_dense_warp_kernel<3>; break;
        case 4: kern = b2p_dense_warp_kernel<4>; break;
        default: kern = b2p_dense_warp_kernel<5>; break;
        }
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // as many one-warp CTAs as fit the SMs at once; each walks the worlds with that stride
        const int per_sm = (int)std::min<size_t>(32, (228 * 1024) / (smem + 1024));
        const int grid = std::max(1, std::min(P->W, std::max(sm_count, 1) * std::max(per_sm, 1)));
        kern<<<grid, 32, smem, stream>>>(*P, *D);
        return cudaGetLastError();
    }
    const int grid = (P->W + DBD - 1) / DBD;
    b2p_dense_lcp_kernel<<<grid, DBD, 0, stream>>>(*P, *D);
    return cudaGetLastError();
}

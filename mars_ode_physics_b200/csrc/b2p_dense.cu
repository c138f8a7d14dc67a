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
//   * ODE updates an LDL^T factor of A(C,C) row by row; here it is a Cholesky factor that grows by one row when an
//     index enters C and is rebuilt when one leaves;
//   * islands: A is block diagonal over a world's islands, so one problem per world has the same solution as one
//     per island, whichever row order (world block / ODE islands) the assemble kernel used.
// Arithmetic in FP64 (ODE's dWorldStep is used with dDOUBLE; a Cholesky factor of a constraint matrix whose only
// regularisation is cfm ~ 1e-5 .. 1e-10 is not an FP32 computation), rows and results FP32 like the rest of the step.
// Per-world scratch (A, the factor, the vectors, the index sets) lives in global memory, world index fastest: a
// warp whose lanes are in step reads full lines.  This is the exact, O(m^3) stepper -- tens of times slower than the
// SOR kernels by construction; it is a correctness path, measured but not tuned.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

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

} // namespace

// per-world scratch of the dense stepper, in elements: doubles of A / of the factor / of the vectors, bytes of the sets
extern "C" void b2p_dense_scratch_sizes(int nb, int maxrows, size_t *mat_doubles, size_t *vec_doubles, size_t *set_bytes)
{
    *mat_doubles = (size_t)maxrows * maxrows;
    *vec_doubles = (size_t)maxrows * 9 + (size_t)(nb + 1) * 6;
    *set_bytes = (size_t)maxrows * 5;
}

extern "C" cudaError_t b2p_launch_dense(const StepParams *P, const DenseScratch *D, cudaStream_t stream)
{
    const int grid = (P->W + DBD - 1) / DBD;
    b2p_dense_lcp_kernel<<<grid, DBD, 0, stream>>>(*P, *D);
    return cudaGetLastError();
}

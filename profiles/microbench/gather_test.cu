// gather_test.cu -- what limits the per-lane row gathers of the SOR kernel's streaming sweeps?
// 148 x 2 CTAs x 256 threads, each thread (= world) fetches 31 rows of 64 bytes in a random per-world order from a
// 130 MB array (the footprint of the rover rows), several access shapes:
//   A  two LDG.256 per row, layout [slot][unit][world][32 B]      (the tree: every lane of a request in its own line)
//   B  two LDG.256 per row, layout [slot][world][64 B]            (a lane's two units in one line: 32 lines/request)
//   C  pair exchange: lanes 2i,2i+1 fetch the two halves of one row per request (16 lines/request), layout B
//   D  four LDG.128 per row, layout A
//   E  cp.async 16 B x 4 per row into shared memory, layout A
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_test gather_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int W = 65536, R = 31, TW = 256;
struct __align__(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const void *p) {
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w),
                 "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
template <int MODE> __global__ void __launch_bounds__(TW, 1) k(const float4 *__restrict__ rows, float *out, int passes, int global_scatter)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    // worlds of this thread: a scattered one inside the CTA's window (like the sorted worlds of the SOR kernel)
    const int base = blockIdx.x * 224;
    // global_scatter: the CTA's worlds come from all over the batch (a globally rank-sorted assignment) instead of
    // one 224-world window
    const int w = global_scatter ? (int)(((unsigned)(base + tid) * 40503u + 12345u) % (unsigned)W) : base + (tid * 37) % 224;
    uint32_t s = 12345u + w;
    float acc = 0.f;
    for (int pass = 0; pass < passes; pass++) {
        uint8_t ord[R];
        for (int i = 0; i < R; i++) ord[i] = i;
        for (int i = R - 1; i > 0; i--) { int j = lcg(s) % (i + 1); uint8_t t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
        if (MODE == 0) {
#pragma unroll 3
            for (int p = 0; p < R; p++) {
                const int sl = ord[p];
                F8 a = ldg256(rows + ((size_t)(sl * 2 + 0) * W + w) * 2), b = ldg256(rows + ((size_t)(sl * 2 + 1) * W + w) * 2);
                acc += a.lo.x + a.hi.w + b.lo.y + b.hi.z;
            }
        } else if (MODE == 1) {
#pragma unroll 3
            for (int p = 0; p < R; p++) {
                const int sl = ord[p];
                const float4 *q = rows + ((size_t)sl * W + w) * 4;
                F8 a = ldg256(q), b = ldg256(q + 2);
                acc += a.lo.x + a.hi.w + b.lo.y + b.hi.z;
            }
        } else if (MODE == 2) {
#pragma unroll 3
            for (int p = 0; p < R; p++) {
                const int sl = ord[p];
                const int wn = __shfl_xor_sync(0xffffffffu, w, 1), sn = __shfl_xor_sync(0xffffffffu, sl, 1);
                const bool odd = lane & 1;
                // request 1: the row of the even lane, request 2: the row of the odd lane; each lane fetches its half
                const int w1 = odd ? wn : w, s1 = odd ? sn : sl, w2 = odd ? w : wn, s2 = odd ? sl : sn;
                F8 a = ldg256(rows + ((size_t)s1 * W + w1) * 4 + (odd ? 2 : 0));
                F8 b = ldg256(rows + ((size_t)s2 * W + w2) * 4 + (odd ? 2 : 0));
                // exchange: even lanes keep a (own first half) and need the partner's a; odd keep b
                F8 give = odd ? a : b, got;
                got.lo.x = __shfl_xor_sync(0xffffffffu, give.lo.x, 1); got.lo.y = __shfl_xor_sync(0xffffffffu, give.lo.y, 1);
                got.lo.z = __shfl_xor_sync(0xffffffffu, give.lo.z, 1); got.lo.w = __shfl_xor_sync(0xffffffffu, give.lo.w, 1);
                got.hi.x = __shfl_xor_sync(0xffffffffu, give.hi.x, 1); got.hi.y = __shfl_xor_sync(0xffffffffu, give.hi.y, 1);
                got.hi.z = __shfl_xor_sync(0xffffffffu, give.hi.z, 1); got.hi.w = __shfl_xor_sync(0xffffffffu, give.hi.w, 1);
                F8 u0 = odd ? got : a, u1 = odd ? b : got;
                acc += u0.lo.x + u0.hi.w + u1.lo.y + u1.hi.z;
            }
        } else if (MODE == 3) {
#pragma unroll 3
            for (int p = 0; p < R; p++) {
                const int sl = ord[p];
                const float4 *q0 = rows + ((size_t)(sl * 2 + 0) * W + w) * 2, *q1 = rows + ((size_t)(sl * 2 + 1) * W + w) * 2;
                float4 a = __ldg(q0), b = __ldg(q0 + 1), c = __ldg(q1), d = __ldg(q1 + 1);
                acc += a.x + b.w + c.y + d.z;
            }
        } else {
            float4 *st = (float4 *)sm + tid; // [R][4][TW]
            for (int p = 0; p < R; p++) {
                const int sl = ord[p];
                const float4 *q0 = rows + ((size_t)(sl * 2 + 0) * W + w) * 2, *q1 = rows + ((size_t)(sl * 2 + 1) * W + w) * 2;
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(st + (size_t)((p % 6) * 4) * TW);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(q0));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + TW * 16), "l"(q0 + 1));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 2 * TW * 16), "l"(q1));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 3 * TW * 16), "l"(q1 + 1));
                asm volatile("cp.async.commit_group;");
                if (p >= 5) {
                    asm volatile("cp.async.wait_group 5;");
                    acc += st[(size_t)(((p - 5) % 6) * 4) * TW].x;
                }
            }
            asm volatile("cp.async.wait_group 0;");
        }
    }
    out[blockIdx.x * TW + tid] = acc;
}
template <int MODE> void run(const char *name, const float4 *rows, float *out, int gs = 0)
{
    const int grid = (W + 223) / 224, passes = 3;
    size_t smem = MODE == 4 ? (size_t)6 * 4 * TW * 16 : 0;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    smem = smem < 120 * 1024 ? 120 * 1024 : smem; // one CTA per SM, as the SOR kernel
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, TW, smem>>>(rows, out, passes, gs);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) k<MODE><<<grid, TW, smem>>>(rows, out, passes, gs);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-50s %8.2f us per pass of %d rows x %d worlds (%s)\n", name, ms * 1000 / 10 / passes, R, W, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float4 *rows; float *out;
    cudaMalloc(&rows, (size_t)(R + 1) * 2 * W * 32);
    cudaMemset(rows, 0, (size_t)(R + 1) * 2 * W * 32);
    cudaMalloc(&out, 4 * 300 * TW);
    run<0>("A LDG.256 x2, [slot][unit][world]", rows, out);
    run<1>("B LDG.256 x2, [slot][world][64B]", rows, out);
    run<2>("C pair exchange (16 lines/request), layout B", rows, out);
    run<3>("D LDG.128 x4, layout A", rows, out);
    run<4>("E cp.async 16B x4 -> smem, layout A", rows, out);
    run<0>("A with the CTA's worlds scattered over the whole batch", rows, out, 1);
    return 0;
}

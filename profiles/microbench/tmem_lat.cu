// scratch: does tcgen05.wait::ld of one warp wait for the other warp sharing its TMEM quadrant / scheduler?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
template <int WORK> __global__ void __launch_bounds__(256) k(long long *cyc, int nwarps_active, int mode)
{
    __shared__ uint32_t s_base;
    __shared__ float sm[256 * 8];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t base = s_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    for (int i = 0; i < 8; i++) sm[i * 256 + threadIdx.x] = 1.0f + i;
    uint32_t v[16];
    float acc = threadIdx.x;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps_active) {
        t0 = clock64();
        for (int it = 0; it < 2000; it++) {
            if (mode == 1) {
                tm_ld16(base + (it & 15) * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                acc += __uint_as_float(v[it & 15] & 0x3f800000);
            }
            if (mode == 3) { // issue only
                tm_ld16(base + (it & 15) * 16, v);
            }
            if (mode == 4) { // wait only (nothing outstanding)
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            }
            if (mode == 5) { // x8 load in the prefetch pattern
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                acc += __uint_as_float(v[it & 7] & 0x3f800000);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(base + (it & 15) * 16) : "memory");
            }
            if (mode == 6) { // two rows per load (x32), wait every other iteration
                if ((it & 1) == 0) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    acc += __uint_as_float(v[it & 15] & 0x3f800000);
                    tm_ld16(base + (it & 15) * 16, v);
                }
            }
            if (mode == 2) { // prefetch pattern: wait for the load issued one iteration ago, consume, issue the next
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                acc += __uint_as_float(v[it & 15] & 0x3f800000);
                tm_ld16(base + (it & 15) * 16, v);
            }
            // dependent chain of WORK smem round trips + FMAs (stands for one row relaxation)
#pragma unroll
            for (int j = 0; j < WORK; j++) {
                float x = sm[(j & 7) * 256 + threadIdx.x];
                acc = acc * 0.999f + x;
                sm[(j & 7) * 256 + threadIdx.x] = acc * 1e-3f;
            }
        }
        t1 = clock64();
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    acc += __uint_as_float(v[3] & 0x3f800000);
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0 + (acc == 12345.f);
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_base), "n"(512));
}
int main()
{
    long long *c; cudaMalloc(&c, 8 * 8);
    long long h[8];
    for (int mode = 0; mode < 7; mode++)
        for (int nw : {1, 8}) {
            cudaMemset(c, 0, 64);
            k<8><<<1, 256>>>(c, nw, mode);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
            printf("mode %d (tmem ld %s) warps %d: cycles/iter", mode, mode == 0 ? "off" : (mode == 1 ? "ld+wait" : (mode == 2 ? "prefetch" : (mode == 3 ? "issue only" : (mode == 4 ? "wait only" : (mode == 5 ? "x8 prefetch" : "every 2nd iter"))))), nw);
            for (int w = 0; w < nw; w++) printf(" %lld", h[w] / 2000);
            printf("  %s\n", cudaGetErrorString(e));
        }
    return 0;
}

// scratch: TMEM as per-thread row store -- st/ld round trip + latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__global__ void __launch_bounds__(128) k(uint32_t *out, long long *cyc)
{
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int r = 0; r < 32; r++) {
        for (int i = 0; i < 16; i++) v[i] = (blockIdx.x << 24) | (threadIdx.x << 16) | (r << 8) | i;
        tm_st16(base + r * 16, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    uint32_t bad = 0;
    long long t0 = clock64();
    for (int r = 0; r < 32; r++) {
        tm_ld16(base + r * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int i = 0; i < 16; i++) bad += v[i] != ((blockIdx.x << 24) | (threadIdx.x << 16) | (r << 8) | i);
    }
    long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = bad;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / 32;
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s_base), "n"(512));
}
int main()
{
    uint32_t *d; long long *c;
    const int nb = 600;
    cudaMalloc(&d, nb * 128 * 4); cudaMalloc(&c, nb * 8);
    k<<<nb, 128>>>(d, c);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    static uint32_t h[600 * 128]; static long long hc[600];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
    long bad = 0; for (int i = 0; i < nb * 128; i++) bad += h[i];
    printf("mismatches %ld, ld+wait+check cycles per row: %lld %lld\n", bad, hc[0], hc[599]);
    return 0;
}

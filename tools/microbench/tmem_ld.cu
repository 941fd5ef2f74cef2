// Microbenchmark: operand fetch of 8 doubles per lane from TMEM (tcgen05.ld.32x32b.x16) vs shared memory (4 x LDS.128),
// each followed by 8 dependent DADDs, at 8..32 warps per SM.  Prints cycles per fetch per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: TMEM, 1: shared, 2: both alternating (2 of 3 fetches from shared memory)
__global__ void __launch_bounds__(1024, 1) k(double *out, int iters, long long *cycles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *xs = reinterpret_cast<double *>(smem);            // 20 columns x 256 doubles
    for (int i = threadIdx.x; i < 20 * 256; i += blockDim.x) xs[i] = 1e-9 * i;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    if (warp < 4) {                                           // one warp per lane quadrant fills its copy of the 20 columns
        for (int c = 0; c < 20; ++c) {
            uint32_t r[16];
            for (int j = 0; j < 4; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(xs + c * 256 + 2 * lane + 64 * j);
                r[4 * j] = __double2loint(v.x); r[4 * j + 1] = __double2hiint(v.x);
                r[4 * j + 2] = __double2loint(v.y); r[4 * j + 3] = __double2hiint(v.y);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(tbase + 16u * c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t xs0 = (uint32_t)__cvta_generic_to_shared(xs) + 16u * lane;
    uint32_t col = (warp * 7) % 20;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        col = col + 7; if (col >= 20) col -= 20;
        const bool use_tmem = MODE == 0 || (MODE == 2 && (it % 3) == 0);
        if (use_tmem) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(tbase + 16u * col) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += __hiloint2double(r[2 * i + 1], r[2 * i]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double x, y;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(xs0 + 2048u * col + 512u * j));
                a[2 * j] += x; a[2 * j + 1] += y;
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16, 24, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32, 48 * 1024>>>(out, iters, cyc);
                if (mode == 1) k<1><<<148, warps * 32, 48 * 1024>>>(out, iters, cyc);
                if (mode == 2) k<2><<<148, warps * 32, 48 * 1024>>>(out, iters, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
            }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            printf("%s warps %2d: %.1f cycles per fetch per warp, %.2f cycles per fetch per SM (2 KB each)\n",
                   mode == 0 ? "tmem  " : mode == 1 ? "shared" : "mixed ", warps, avg / iters, avg / iters / warps);
        }
    return 0;
}

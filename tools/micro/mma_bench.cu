// Microbenchmark: cycles per tcgen05.mma (kind::f16, cta_group::1) as a function of tile shape,
// shared-memory layout (no-swizzle canonical vs 128B swizzle) and number of independent accumulators.
// Descriptors are precomputed; the timed loop is 4 unrolled MMAs per iteration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../zaremba_b200/csrc/tc_common.cuh"
using namespace zrb::tc;

template <int NACC>
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int swz, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(M, N, 0, 0);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
        const int GA = M / 8, GBn = N / 8;
        const uint32_t stride = (N + 31) / 32 * 32;
        uint64_t da[4], db[4];
        for (int k = 0; k < 4; ++k) {
            if (swz) {
                da[k] = make_smem_desc(a_addr + k * 32, 16, 1024, kSwizzle128B);
                db[k] = make_smem_desc(b_addr + k * 32, 16, 1024, kSwizzle128B);
            } else {
                da[k] = make_smem_desc(a_addr + k * 2 * GA * 128, GA * 128, 128, kSwizzleNone);
                db[k] = make_smem_desc(b_addr + k * 2 * GBn * 128, GBn * 128, 128, kSwizzleNone);
            }
        }
        for (int j = 0; j < 4; ++j) umma_f16(tm + (j % NACC) * stride, da[j], db[j], idesc, 0u);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16(tm + (j % NACC) * stride, da[j], db[j], idesc, 1u);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("M,N,swizzle,accs,iters,issue_clk_per_mma,total_clk_per_mma\n");
    int Ms[] = {64, 128};
    int Ns[] = {8, 16, 24, 32, 48, 64, 128, 256};
    const int iters = 4096;
    for (int swz = 0; swz < 2; ++swz) for (int mi = 0; mi < 2; ++mi) for (int ni = 0; ni < 8; ++ni) for (int accs = 1; accs <= 4; accs *= 4) {
        int M = Ms[mi], N = Ns[ni];
        if (M == 128 && N % 16) continue;
        if (accs * ((N + 31) / 32 * 32) > 512) continue;
        if (accs == 1) bench<1><<<1, 128, 190 * 1024>>>(M, N, swz, iters, d);
        else bench<4><<<1, 128, 190 * 1024>>>(M, N, swz, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%d,%d,%d,%d,ERR %s\n", M, N, swz, accs, cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%d,%d,%d,%d,%d,%.1f,%.1f\n", M, N, swz, accs, iters, (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}

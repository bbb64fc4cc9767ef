// Microbenchmark: cost of draining TMEM accumulators with different tcgen05.ld shapes, and the register
// fragment layout of the 16-lane shapes (an M=64 accumulator only occupies lanes 0-15 of each 32-lane quadrant,
// so a 32x32b load moves 50% padding).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../zaremba_b200/csrc/tc_common.cuh"
using namespace zrb::tc;

__device__ __forceinline__ void st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_16x256(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x128x2(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x64x4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// mode 0: 32x32b.x8   1: 16x256b.x1   2: 16x128b.x2   3: 16x64b.x4      (all: 8 columns per instruction)
template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(int iters, int inflight, long long* out, uint32_t* dump) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    if (warp < 4) {   // lane L, column c := 1000 * L + c
        for (int c0 = 0; c0 < 512; c0 += 8) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) v[i] = 1000u * (32 * warp + lane) + c0 + i;
            st_32x8(tm + lane_base + c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (dump && warp == 1) {   // fragment layout of quadrant 1, columns 16..23
        uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (MODE == 0) tmem_ld_32x8(tm + lane_base + 16, v);
        else if (MODE == 1) ld_16x256(tm + lane_base + 16, (uint32_t(&)[4])v);
        else if (MODE == 2) ld_16x128x2(tm + lane_base + 16, (uint32_t(&)[4])v);
        else ld_16x64x4(tm + lane_base + 16, (uint32_t(&)[4])v);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) dump[lane * 8 + i] = v[i];
    }
    __syncthreads();
    uint32_t sink = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < inflight; ++k) {
            uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const uint32_t addr = tm + lane_base + ((it * inflight + k) * 8 & 255) + (warp >> 2) * 256;
            if (MODE == 0) tmem_ld_32x8(addr, v);
            else if (MODE == 1) ld_16x256(addr, (uint32_t(&)[4])v);
            else if (MODE == 2) ld_16x128x2(addr, (uint32_t(&)[4])v);
            else ld_16x64x4(addr, (uint32_t(&)[4])v);
            sink ^= v[0] ^ v[3] ^ v[7];
        }
        tmem_ld_wait();
    }
    long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (sink == 0x12345u) out[9] = sink;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int MODE>
void run(const char* name, long long* d_out, uint32_t* d_dump) {
    uint32_t h[256];
    bench<MODE><<<1, 128>>>(1, 1, d_out, d_dump);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d_dump, sizeof(h), cudaMemcpyDeviceToHost);
    printf("# layout %s (quadrant 1, columns 16..23): thread -> (lane,col) per register\n", name);
    for (int t = 0; t < 32; t += 1) {
        printf("#  t%02d:", t);
        for (int i = 0; i < (MODE == 0 ? 8 : 4); ++i) printf(" (%u,%u)", h[t * 8 + i] / 1000, h[t * 8 + i] % 1000);
        printf("\n");
    }
    for (int warps : {4, 8}) {
        for (int inflight : {1, 3, 6, 12}) {
            const int iters = 2000;
            bench<MODE><<<1, warps * 32>>>(iters, inflight, d_out, nullptr);
            cudaError_t e = cudaDeviceSynchronize();
            long long o[8];
            cudaMemcpy(o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < warps; ++w) mx = o[w] > mx ? o[w] : mx;
            // CTA-wide: warps*inflight instructions of 8 columns per iteration
            printf("%s,%d,%d,%.1f,%.2f,%s\n", name, warps, inflight, (double)mx / iters,
                   (double)mx / iters / (warps * inflight), cudaGetErrorString(e));
        }
    }
}

int main() {
    long long* d_out; uint32_t* d_dump;
    cudaMalloc(&d_out, 16 * sizeof(long long));
    cudaMalloc(&d_dump, 256 * 4);
    printf("shape,warps,inflight,clk_per_iter,clk_per_ld_cta,status\n");
    run<0>("32x32b.x8", d_out, d_dump);
    run<1>("16x256b.x1", d_out, d_dump);
    run<2>("16x128b.x2", d_out, d_dump);
    run<3>("16x64b.x4", d_out, d_dump);
    return 0;
}

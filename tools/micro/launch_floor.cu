// What does an EMPTY kernel with the recurrence kernels' launch shape cost?  (126-128 CTAs x 416 threads, ~210 KB dynamic
// shared memory each, plain / cooperative / cluster launches), back to back on one stream, CUDA-event timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o launch_floor_bench launch_floor.cu && ./launch_floor_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void empty_kernel(int* p) { extern __shared__ char smem[]; if (threadIdx.x == 0 && p) smem[0] = 1; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    CK(cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct V { const char* name; int grid, smem, cluster, coop; } vs[] = {
        {"plain, 8 KB smem", 126, 8 * 1024, 0, 0}, {"plain, 210 KB smem", 126, 210 * 1024, 0, 0},
        {"cooperative, 210 KB", 126, 210 * 1024, 0, 1}, {"cluster 2 + cooperative, 210 KB", 126, 210 * 1024, 2, 1},
        {"cluster 8 + cooperative, 215 KB", 120, 215 * 1024, 8, 1}, {"cluster 8, 215 KB", 120, 215 * 1024, 8, 0}};
    for (auto& v : vs) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(v.grid); cfg.blockDim = dim3(416); cfg.dynamicSmemBytes = v.smem; cfg.stream = s;
        cudaLaunchAttribute at[2]; int na = 0;
        if (v.cluster) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = v.cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
        if (v.coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
        cfg.attrs = at; cfg.numAttrs = na;
        for (int i = 0; i < 20; ++i) CK(cudaLaunchKernelEx(&cfg, empty_kernel, (int*)nullptr));
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < 200; ++i) CK(cudaLaunchKernelEx(&cfg, empty_kernel, (int*)nullptr));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-36s %6.2f us per back-to-back launch\n", v.name, ms * 1000 / 200);
    }
    return 0;
}

// Does alternating between a ~210 KB-shared-memory kernel and a no-shared-memory kernel cost an SM
// shared-memory/L1 re-partition per switch?  Pairs (big, small) back to back, CUDA-event timed, with the small
// kernel's preferred carve-out left at the default and then forced to "max shared".
//   nvcc -gencode arch=compute_100a,code=sm_100a -o carveout_switch_bench carveout_switch.cu && ./carveout_switch_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void big_kernel(int* p) { extern __shared__ char smem[]; if (threadIdx.x == 0 && p) smem[0] = 1; }
__global__ void small_kernel(float* x, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) x[i] += 1.f; }
__global__ void small_kernel2(float* x, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) x[i] += 1.f; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    CK(cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024));
    CK(cudaFuncSetAttribute(small_kernel2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    float* x; const int n = 148 * 256 * 4; CK(cudaMalloc(&x, n * 4)); CK(cudaMemset(x, 0, n * 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 4; ++mode) {
        const char* names[] = {"small only (default carve-out)", "big only", "big + small (default carve-out)", "big + small (max-shared carve-out)"};
        auto once = [&]() {
            if (mode != 0) big_kernel<<<126, 416, 210 * 1024, s>>>(nullptr);
            if (mode == 0 || mode == 2) small_kernel<<<n / 256, 256, 0, s>>>(x, n);
            if (mode == 3) small_kernel2<<<n / 256, 256, 0, s>>>(x, n);
        };
        for (int i = 0; i < 20; ++i) once();
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < 200; ++i) once();
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-40s %6.2f us per iteration\n", names[mode], ms * 1000 / 200);
    }
    return 0;
}

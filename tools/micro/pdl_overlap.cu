// Does a programmatic dependent launch start on the SMs a persistent 128-CTA kernel leaves idle, and for which launch
// flavours of the primary (plain, cooperative, cluster, cluster + cooperative)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_overlap pdl_overlap.cu && ./pdl_overlap
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// primary: every CTA triggers its dependents at once, then spins for `ns` nanoseconds holding ~200 KB of smem
__global__ void primary(unsigned long long ns, unsigned long long* stamps, int trigger) {
    extern __shared__ char smem[];
    if (threadIdx.x == 0) {
        smem[0] = 1;
        unsigned long long t0 = gtime();
        if (blockIdx.x == 0) stamps[0] = t0;
        if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        while (gtime() - t0 < ns) {}
        if (blockIdx.x == 0) stamps[1] = gtime();
    }
    __syncthreads();
}

// secondary: records when each CTA started; waits for the primary before exiting
__global__ void secondary(unsigned long long* starts, unsigned long long work_ns, int tail_wait) {
    extern __shared__ char smem[];
    if (threadIdx.x == 0) {
        smem[0] = 1;
        unsigned long long t0 = gtime();
        starts[blockIdx.x] = t0;
        while (gtime() - t0 < work_ns) {}
        if (tail_wait == 1 || (tail_wait == 2 && blockIdx.x == 0)) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    __syncthreads();
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
    const int smem = 200 * 1024, nprim = 128, nsec = 148;
    CK(cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long *stamps, *starts, h_stamps[2], h_starts[nsec];
    CK(cudaMalloc(&stamps, 16)); CK(cudaMalloc(&starts, nsec * 8));
    cudaStream_t s_created; CK(cudaStreamCreate(&s_created));
    const char* names[] = {"plain", "cooperative", "cluster4", "cluster4+cooperative"};
    for (int which_stream = 0; which_stream < 3; ++which_stream)
    for (int trigger = 1; trigger >= 0; --trigger)
    for (int flavour = 0; flavour < 4; ++flavour) {
        if (which_stream >= 1) { if (trigger == 0 || (flavour != 0 && flavour != 3)) continue; }
        const int tail_mode = which_stream == 2 ? 2 : 1;   // 2: only CTA 0 of the dependent grid waits for the primary
        cudaStream_t s = which_stream == 0 ? s_created : (cudaStream_t)0;   // 1: the legacy default stream
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemsetAsync(starts, 0, nsec * 8, s));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nprim); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = s;
            cudaLaunchAttribute at[2]; int na = 0;
            if (flavour >= 2) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = 4; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
            if (flavour == 1 || flavour == 3) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
            cfg.attrs = at; cfg.numAttrs = na;
            CK(cudaLaunchKernelEx(&cfg, primary, 200000ull, stamps, trigger));
            cudaLaunchConfig_t c2 = {};
            c2.gridDim = dim3(nsec); c2.blockDim = dim3(128); c2.dynamicSmemBytes = smem; c2.stream = s;
            cudaLaunchAttribute a2[1];
            a2[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a2[0].val.programmaticStreamSerializationAllowed = 1;
            c2.attrs = a2; c2.numAttrs = 1;
            CK(cudaLaunchKernelEx(&c2, secondary, starts, 20000ull, tail_mode));
            CK(cudaStreamSynchronize(s));
            CK(cudaMemcpy(h_stamps, stamps, 16, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_starts, starts, nsec * 8, cudaMemcpyDeviceToHost));
            int early = 0; long long first = 1ll << 62;
            for (int i = 0; i < nsec; ++i) {
                if (h_starts[i] < h_stamps[1]) ++early;
                long long d = (long long)h_starts[i] - (long long)h_stamps[0];
                if (d < first) first = d;
            }
            if (rep == 1)
                printf("%s primary %-22s trigger=%d: primary ran %.1f us; %3d of %d dependent CTAs started before it ended; first dependent CTA at +%.1f us\n",
                       which_stream == 2 ? "[legacy stream, only CTA 0 waits]" : which_stream ? "[legacy default stream]" : "[created stream]", names[flavour], trigger, (h_stamps[1] - h_stamps[0]) / 1e3, early, nsec, first / 1e3);
        }
    }
    return 0;
}

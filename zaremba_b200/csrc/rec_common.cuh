// Helpers shared by the persistent recurrence kernels (forward / backward).
#pragma once
#include "tc_common.cuh"
#include "tc_kernels.h"

namespace zrb {
using namespace tc;

constexpr int kRecThreads = 192;  // warps 0-3 epilogue, warp 4 MMA + TMEM, warp 5 loader
constexpr long long kSpinCycles = 6000000000ll;  // ~3 s at 2 GHz: a lost wake-up traps instead of hanging the GPU

__device__ __forceinline__ void bounded_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t n = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++n & 0xFFFu) == 0) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kSpinCycles) asm volatile("trap;");
        }
    }
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}


// spin on a global counter (grid barrier) with acquire semantics and the same bounded wait
__device__ __forceinline__ void grid_counter_wait(const unsigned int* counter, unsigned int target) {
    uint32_t n = 0;
    long long t0 = 0;
    while (ld_acquire_gpu(counter) < target) {
        if ((++n & 0x3FFu) == 0) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kSpinCycles) asm volatile("trap;");
        }
    }
}

}  // namespace zrb

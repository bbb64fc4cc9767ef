// Shared host/device helpers for libzaremba_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/zaremba_b200.h"

namespace zrb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_live_tc_ctx[64];   // live tcgen05-engine contexts per device (index = device & 63)
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define ZRB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            zrb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                           cudaGetErrorString(e_));                                      \
            return ZRB_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

#define ZRB_KERNEL_CHECK()                                                               \
    do {                                                                                 \
        zrb::count_launch();                                                             \
        ZRB_CUDA(cudaGetLastError());                                                    \
    } while (0)

#define ZRB_REQUIRE(cond, ...)                                                           \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            zrb::set_error(__VA_ARGS__);                                                 \
            return ZRB_E_INVALID;                                                        \
        }                                                                                \
    } while (0)

#define ZRB_TRY(expr)                                                                    \
    do {                                                                                 \
        int rc_ = (expr);                                                                \
        if (rc_ != ZRB_OK) return rc_;                                                   \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter-based: dropout keep-masks are a pure
// function of (seed, step, site, element), so backward regenerates them instead of
// storing or re-reading a mask tensor.
// ---------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// How the keep-mask of one dropout site is obtained.
struct MaskSrc {
    const uint8_t* explicit_mask;  // if non-null: [n] bytes, 1 = keep (replayed reference masks)
    uint32_t k0, k1;               // Philox key   = seed
    uint32_t c2, c3;               // Philox ctr hi = (site, step)
    uint32_t thresh;               // keep iff (r >> 8) >= thresh, thresh = round(p * 2^24)
    float scale;                   // 1 / (1 - p)
    int active;                    // 0: identity (eval mode or p == 0)
};

__host__ inline MaskSrc make_mask_src(const uint8_t* explicit_mask, uint64_t seed, uint64_t step, int site,
                                      float p, int train) {
    MaskSrc m;
    m.explicit_mask = explicit_mask;
    m.k0 = (uint32_t)seed;
    m.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
    m.c2 = (uint32_t)site;
    m.c3 = (uint32_t)step;
    double t = (double)p * 16777216.0;
    m.thresh = (uint32_t)(t + 0.5);
    m.scale = (float)(1.0 / (1.0 - (double)p));
    m.active = (train && p > 0.f) ? 1 : 0;
    return m;
}

// keep flags of the 4 consecutive elements [4*g, 4*g+3] packed in bits 0..3
__device__ inline uint32_t mask_keep4(const MaskSrc& m, uint64_t g, uint64_t n_total) {
    if (m.explicit_mask) {
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint64_t e = 4 * g + i;
            if (e < n_total && m.explicit_mask[e]) bits |= 1u << i;
        }
        return bits;
    }
    Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), m.c2, m.c3, m.k0, m.k1);
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if ((r.v[i] >> 8) >= m.thresh) bits |= 1u << i;
    return bits;
}

// multiplier (0 or scale) for a single element e
__device__ inline float mask_mul1(const MaskSrc& m, uint64_t e, uint64_t n_total) {
    if (!m.active) return 1.f;
    uint32_t bits = mask_keep4(m, e >> 2, n_total);
    return ((bits >> (e & 3)) & 1u) ? m.scale : 0.f;
}

__device__ inline float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

__device__ inline float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace zrb

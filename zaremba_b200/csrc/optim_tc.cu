// SGD update fused with the rebuild of the fp16 operand images (tensor-core engine).
// The update pass already holds every new weight in registers; writing its fp16 images from there
// removes the separate pack pass (which re-read 348 MB of fp32 weights per step at the Large config).
#include "tc_kernels.h"

namespace zrb {

struct PackSpec {
    __half* row_img;     // [rows, ld] row-major image or null
    int64_t ld;
    __half* fwd_img;     // recurrent forward slices  [cta][kc][g][8][8] or null
    int fU, fG, fKc;
    __half* bwd_img;     // recurrent backward slices [cluster][4][kc][g][8][8] or null
    int bUC, bG, bKc;
};

// matrix [rows, cols] (cols % 4 == 0): g *= coef; p -= lr*g; images of the new p.
// For the recurrent images rows = 4H (gate q = row / H, unit j = row % H), cols = H.
__global__ void update_pack_kernel(float* __restrict__ p, float* __restrict__ g, int rows, int cols, float lr,
                                   const float* __restrict__ scalars, PackSpec sp) {
    const float coef = scalars[1];
    const int c4 = cols >> 2;
    const int64_t total = (int64_t)rows * c4;
    const int H = cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / c4), c = (int)(i % c4) << 2;
        float4* g4 = reinterpret_cast<float4*>(g + (int64_t)r * cols + c);
        float4* p4 = reinterpret_cast<float4*>(p + (int64_t)r * cols + c);
        float4 gv = __ldcs(g4), pv = __ldcs(p4);
        gv.x *= coef; gv.y *= coef; gv.z *= coef; gv.w *= coef;
        pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
        __stcs(g4, gv);
        __stcs(p4, pv);
        const __half h0 = __float2half_rn(pv.x), h1 = __float2half_rn(pv.y), h2 = __float2half_rn(pv.z),
                     h3 = __float2half_rn(pv.w);
        if (sp.row_img) {
            __half2* d = reinterpret_cast<__half2*>(sp.row_img + (int64_t)r * sp.ld + c);
            d[0] = __halves2half2(h0, h1);
            d[1] = __halves2half2(h2, h3);
        }
        if (sp.fwd_img) {   // W_hh[q*H + j, k..k+3] -> slice of the CTA owning unit j, row 4u+q, K chunk k/8
            const int q = r / H, j = r % H, k = c;
            const int cta = j / sp.fU, u = j % sp.fU, row = 4 * u + q;
            const int64_t idx = (((int64_t)cta * sp.fKc + (k >> 3)) * sp.fG + (row >> 3)) * 64 + (row & 7) * 8 + (k & 7);
            __half2* d = reinterpret_cast<__half2*>(sp.fwd_img + idx);
            d[0] = __halves2half2(h0, h1);
            d[1] = __halves2half2(h2, h3);
        }
        if (sp.bwd_img) {   // W_hh[q*H + j, units c..c+3] -> cluster owning those units, rank q, K index j
            const int q = r / H, j = r % H;
            const int cl = c / sp.bUC, u = c % sp.bUC;
            const int64_t base = ((((int64_t)cl * 4 + q) * sp.bKc + (j >> 3)) * sp.bG) * 64 + (j & 7);
            sp.bwd_img[base + ((u + 0) >> 3) * 64 + ((u + 0) & 7) * 8] = h0;
            sp.bwd_img[base + ((u + 1) >> 3) * 64 + ((u + 1) & 7) * 8] = h1;
            sp.bwd_img[base + ((u + 2) >> 3) * 64 + ((u + 2) & 7) * 8] = h2;
            sp.bwd_img[base + ((u + 3) >> 3) * 64 + ((u + 3) & 7) * 8] = h3;
        }
    }
}

// W_hh [4H, H] with both recurrent images: a thread owns an 8-row x 4-column tile (rows j..j+7 of one gate
// block), so the backward image -- whose 16-byte vectors hold 8 consecutive K indices (= rows) of one unit
// (= column) -- is written with full 16-byte stores instead of 2-byte scatters.
__global__ void update_pack_whh_kernel(float* __restrict__ p, float* __restrict__ g, int H, float lr,
                                       const float* __restrict__ scalars, PackSpec sp) {
    const float coef = scalars[1];
    const int c4 = H >> 2, jb_n = (H + 7) >> 3;
    const int64_t total = (int64_t)4 * jb_n * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) << 2;
        const int jb = (int)((i / c4) % jb_n), q = (int)(i / ((int64_t)c4 * jb_n));
        const int j0 = jb << 3;
        __half hv[8][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = j0 + e;
            if (j < H) {
                const int64_t off = ((int64_t)q * H + j) * H + c;
                float4* g4 = reinterpret_cast<float4*>(g + off);
                float4* p4 = reinterpret_cast<float4*>(p + off);
                float4 gv = __ldcs(g4), pv = __ldcs(p4);
                gv.x *= coef; gv.y *= coef; gv.z *= coef; gv.w *= coef;
                pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
                __stcs(g4, gv);
                __stcs(p4, pv);
                hv[e][0] = __float2half_rn(pv.x); hv[e][1] = __float2half_rn(pv.y);
                hv[e][2] = __float2half_rn(pv.z); hv[e][3] = __float2half_rn(pv.w);
                if (sp.row_img) {
                    __half2* d = reinterpret_cast<__half2*>(sp.row_img + ((int64_t)q * H + j) * sp.ld + c);
                    d[0] = __halves2half2(hv[e][0], hv[e][1]);
                    d[1] = __halves2half2(hv[e][2], hv[e][3]);
                }
                if (sp.fwd_img) {   // slice of the CTA owning unit j, row 4u+q, K chunk c/8, 4 consecutive K
                    const int cta = j / sp.fU, u = j % sp.fU, row = 4 * u + q;
                    const int64_t idx = (((int64_t)cta * sp.fKc + (c >> 3)) * sp.fG + (row >> 3)) * 64 + (row & 7) * 8 + (c & 7);
                    __half2* d = reinterpret_cast<__half2*>(sp.fwd_img + idx);
                    d[0] = __halves2half2(hv[e][0], hv[e][1]);
                    d[1] = __halves2half2(hv[e][2], hv[e][3]);
                }
            } else {
                hv[e][0] = hv[e][1] = hv[e][2] = hv[e][3] = __float2half_rn(0.f);
            }
        }
        if (sp.bwd_img) {   // units c..c+3 of cluster c/UC, rank q, K chunk jb: one 16-byte vector per unit
            const int cl = c / sp.bUC, u = c % sp.bUC;
            const int64_t base = ((((int64_t)cl * 4 + q) * sp.bKc + jb) * sp.bG) * 64;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                uint4 v;
                v.x = (uint32_t)__half_as_ushort(hv[0][x]) | ((uint32_t)__half_as_ushort(hv[1][x]) << 16);
                v.y = (uint32_t)__half_as_ushort(hv[2][x]) | ((uint32_t)__half_as_ushort(hv[3][x]) << 16);
                v.z = (uint32_t)__half_as_ushort(hv[4][x]) | ((uint32_t)__half_as_ushort(hv[5][x]) << 16);
                v.w = (uint32_t)__half_as_ushort(hv[6][x]) | ((uint32_t)__half_as_ushort(hv[7][x]) << 16);
                *reinterpret_cast<uint4*>(sp.bwd_img + base + ((u + x) >> 3) * 64 + ((u + x) & 7) * 8) = v;
            }
        }
    }
}

int update_pack(float* p, float* g, int rows, int cols, float lr, const float* scalars, __half* row_img, int64_t ld,
                __half* fwd_img, const RecPlan* fp, __half* bwd_img, const RecPlan* bp, cudaStream_t s) {
    PackSpec sp;
    sp.row_img = row_img; sp.ld = ld;
    sp.fwd_img = fwd_img; sp.fU = fp ? fp->U : 1; sp.fG = fp ? fp->G : 1; sp.fKc = fp ? fp->Kc : 1;
    sp.bwd_img = bwd_img; sp.bUC = bp ? 4 * bp->U : 4; sp.bG = bp ? bp->G : 1; sp.bKc = bp ? bp->Kc : 1;
    if (bwd_img && rows == 4 * cols) {
        int64_t total = (int64_t)4 * ((cols + 7) / 8) * (cols / 4);
        int blocks = (int)((total + 127) / 128);
        if (blocks > 148 * 16) blocks = 148 * 16;
        update_pack_whh_kernel<<<blocks, 128, 0, s>>>(p, g, cols, lr, scalars, sp);
        ZRB_KERNEL_CHECK();
        return ZRB_OK;
    }
    int64_t total = (int64_t)rows * (cols / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    update_pack_kernel<<<blocks, 256, 0, s>>>(p, g, rows, cols, lr, scalars, sp);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

}  // namespace zrb

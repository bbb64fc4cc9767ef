// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     C[M,N] (fp32) = alpha * A * B^T (+ bias[N]) (+ C)
// A is [M,K] (K-major) or stored [K,M] (MN-major); B is [N,K] (K-major) or stored [K,N]
// (MN-major).  fp16 operands, fp32 accumulation in TMEM.
//
//   warp 0   TMA producer   : cp.async.bulk.tensor 2-D boxes, 128B swizzle, 6-stage mbarrier ring
//   warp 1   MMA issuer     : one thread issues tcgen05.mma 128x128x16 (cta_group::1), commits
//                             to the ring's "empty" barriers and to the accumulator "full" barrier
//   warp 2   TMEM allocator : two accumulator stages (epilogue of tile i overlaps the MMAs of tile i+1), or, for
//                             the 256x256 tile (MT = 2: two 128-row accumulators that share every B tile), one
//                             stage filling all 512 columns
//   warps 4-11 epilogue     : tcgen05.ld 32x32b, alpha/bias, fp32 stores (two warps per TMEM lane quadrant)
//
// Every batched contraction of the path runs here: X*W_ih^T, the vocabulary projection, their
// dgrads (weights read MN-major from the same fp16 image, no transposed copies) and the
// wgrads (both operands MN-major: contraction over tokens).  Roofline: tensor pipe
// (2*M*N*K flop per call); operands stream once from HBM/L2 via TMA.
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"
#include "tc_host.h"

namespace zrb {

using namespace tc;

constexpr int GBM = 128, GBK = 64;
constexpr int kASubBytes = GBM * GBK * 2;   // one 128-row A sub-tile of a K block
constexpr int kEpiWarps = 8;
constexpr int kEpiBytes = kEpiWarps * 32 * 33 * 4;   // per-epilogue-warp 32x32 transpose tile (padded)
constexpr int kGemmThreads = (4 + kEpiWarps) * 32;
// Tile N is a template parameter: 128 (6 stages) or 256 (4 stages).  The kernel is L2->SM bandwidth bound
// (128x128 tiles pull 32 KB per 2.1 MFLOP K block); 128x256 tiles pull 25% fewer bytes per flop, and their
// 128-clk MMAs hide the issue latency, so 256 is used whenever it still fills the machine.
// MT = 128-row accumulators per CTA tile.  The kernel is L2->SM bound: per K block a 128x256 tile pulls 48 KB for
// 4.2 MFLOP, a 256x256 tile (MT = 2) 64 KB for 8.4 MFLOP -- a third less traffic per flop.  Its two accumulators fill
// TMEM, so the epilogue no longer overlaps the next tile; the shapes of this path give (about) one wave of such
// tiles anyway (e.g. [700 x 10000]: 120 tiles, [6000 x 1500]: 144 tiles on 148 SMs).
template <int GBN, int MT> struct GemmCfg {
    static constexpr int kStages = MT == 2 ? 3 : (GBN == 256 ? 4 : 6);
    static constexpr int kABytes = MT * kASubBytes;
    static constexpr int kBBytes = GBN * GBK * 2;
    static constexpr int kAccStages = MT * GBN >= 512 ? 1 : 2;
    static constexpr int kTmemCols = kAccStages * MT * GBN;
    static constexpr int kSmem = kStages * (kABytes + kBBytes) + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmArgs {
    int M, N, K;
    float alpha;
    const float* bias;
    const float* bias2;   // or null: a second bias vector added with the first (b_ih + b_hh, model.py:35)
    float* C;
    int64_t ldc;
    int accumulate;
    int tiles_m, tiles_n;
    int splits;      // split-K factor (1 or 2); with 2 the epilogue adds atomically into a zeroed C
    float* sumsq_out; // or null: slot [tile * 8 + w] = sum of squares of the outputs epilogue warp w stored for `tile`
                      // (the wgrads feed clip_grad_norm_ from here instead of re-reading 200 MB of gradients)
    float* C2;        // dual launch (or null): a second problem with the same A, shapes and pitches but its own B (tma_b2),
    float* sumsq_out2;  // output and sum-of-squares slots; work items [num_tiles, 2*num_tiles) belong to it (splits == 1)
    const __half* a_tiled;   // EXPERIMENT (zrb_gemm_f16_tiled): pre-tiled, pre-swizzled K-major images ([K block][128-row
    const __half* b_tiled;   // tile][128][64] halves, chunk c of row r stored at c ^ (r % 8)): operand tiles are fetched with
    int a_nt128, b_nt128;    // 1-D bulk copies instead of 2-D tensor loads
    int pdl_trigger;  // release a programmatic dependent enqueued behind this kernel (a recurrence kernel) at once
    int pdl_tail;     // launched as a programmatic dependent of the kernel before it in the stream (it started while that
                      // kernel was still running and consumes none of its outputs): wait for that kernel before exiting,
                      // so that "this grid completed" keeps implying "everything before it in the stream completed"
};

template <bool A_MN, bool B_MN, int GBN, int MT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_b2, GemmArgs p) {
    using Cfg = GemmCfg<GBN, MT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kABytes = Cfg::kABytes;
    constexpr int kBBytes = Cfg::kBBytes;
    constexpr int kAccStages = Cfg::kAccStages;
    constexpr int TM = GBM * MT;   // tile rows
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kABytes;
    float* sEpi = (float*)(smem + kStages * (kABytes + kBBytes));
    uint64_t* bars = (uint64_t*)(smem + kStages * (kABytes + kBBytes) + kEpiBytes);
    uint64_t* full = bars;                       // [kStages]
    uint64_t* empty = bars + kStages;            // [kStages]
    uint64_t* acc_full = bars + 2 * kStages;     // [kAccStages]
    uint64_t* acc_empty = acc_full + kAccStages; // [kAccStages]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + kAccStages);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int num_kb = (p.K + GBK - 1) / GBK;
    const bool dual = p.C2 != nullptr;
    const int num_work = dual ? 2 * num_tiles : num_tiles * p.splits;   // work item w: tile = w % num_tiles; w / num_tiles =
    const int kb_per = (num_kb + p.splits - 1) / p.splits;              //   K range (split-K) or problem (dual launch)

    if (threadIdx.x == 0) {
        if (p.pdl_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        if (dual) tma_prefetch_desc(&tma_b2);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int s = 0; uint32_t ph = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int tile = w % num_tiles, sp = dual ? 0 : w / num_tiles;
            const int kb0 = sp * kb_per, kb1 = min(num_kb, kb0 + kb_per);
            const CUtensorMap* tmb = (dual && w >= num_tiles) ? &tma_b2 : &tma_b;
            const int m0 = (tile % p.tiles_m) * TM, n0 = (tile / p.tiles_m) * GBN;
            const int mt_n = (MT == 2 && m0 + GBM < p.M) ? 2 : 1;   // 128-row sub-tiles that hold real rows
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], mt_n * kASubBytes + kBBytes);
                uint8_t* a = sA + s * kABytes;
                uint8_t* b = sB + s * kBBytes;
                for (int mt = 0; mt < mt_n; ++mt) {
                    if (!A_MN && p.a_tiled) {
                        bulk_load_1d(a + mt * kASubBytes, p.a_tiled + ((size_t)kb * p.a_nt128 + ((m0 + mt * GBM) >> 7)) * (128 * GBK),
                                     kASubBytes, &full[s]);
                    } else if (!A_MN) {
                        tma_load_2d(a + mt * kASubBytes, &tma_a, &full[s], kb * GBK, m0 + mt * GBM);
                    } else {
                        tma_load_2d(a + mt * kASubBytes, &tma_a, &full[s], m0 + mt * GBM, kb * GBK);
                        tma_load_2d(a + mt * kASubBytes + kASubBytes / 2, &tma_a, &full[s], m0 + mt * GBM + 64, kb * GBK);
                    }
                }
                if (!B_MN && p.b_tiled) {
                    bulk_load_1d(b, p.b_tiled + ((size_t)kb * p.b_nt128 + (n0 >> 7)) * (128 * GBK), kBBytes, &full[s]);
                } else if (!B_MN) {
                    tma_load_2d(b, tmb, &full[s], kb * GBK, n0);
                } else {
#pragma unroll
                    for (int j = 0; j < GBN / 64; ++j)
                        tma_load_2d(b + j * (GBK * 128), tmb, &full[s], n0 + 64 * j, kb * GBK);
                }
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks tiles / K blocks and waits on the barriers, so stage indices, phases and
        // the shared-memory descriptors are warp-uniform (they live in uniform registers); one elected
        // lane per K block issues the four tcgen05.mma and the commit.
        constexpr uint32_t idesc = make_idesc_f16(GBM, GBN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int kb0 = (dual ? 0 : w / num_tiles) * kb_per, kb1 = min(num_kb, kb0 + kb_per);
            const int m0 = ((w % num_tiles) % p.tiles_m) * TM;
            const int mt_n = (MT == 2 && m0 + GBM < p.M) ? 2 : 1;
            mbar_wait(&acc_empty[as], aph ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + as * MT * GBN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[s], ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * kABytes), b_addr = smem_u32(sB + s * kBBytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < GBK / 16; ++k) {
                        // K-major, 128B swizzle: rows 128 B apart, 8-row groups 1024 B apart, +32 B per K=16 step.
                        // MN-major, 128B swizzle: 64-element column blocks (BK*128 B apart = LBO), 8 k-rows per
                        // 1024 B group (SBO), +16 k-rows = 2048 B per step.
                        uint64_t db = B_MN ? make_smem_desc(b_addr + k * 2048, GBK * 128, 1024, kSwizzle128B)
                                           : make_smem_desc(b_addr + k * 32, 16, 1024, kSwizzle128B);
                        for (int mt = 0; mt < mt_n; ++mt) {   // the sub-tiles' MMAs share the B descriptor
                            const uint32_t am = a_addr + mt * kASubBytes;
                            uint64_t da = A_MN ? make_smem_desc(am + k * 2048, kASubBytes / 2, 1024, kSwizzle128B)
                                               : make_smem_desc(am + k * 32, 16, 1024, kSwizzle128B);
                            umma_f16(d_tmem + mt * GBN, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (kb == kb1 - 1) umma_commit(&acc_full[as]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
            if (kb1 <= kb0 && elect_one()) umma_commit(&acc_full[as]);   // empty K range: nothing to wait for
            __syncwarp();
            if (++as == kAccStages) { as = 0; aph ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps =====================
        // tcgen05.ld hands thread i of the warp accumulator row 32q+i; a padded 32x32 shared-memory transpose turns
        // that into row-contiguous 128-byte global stores (one row per instruction).  Two warps share each TMEM lane
        // quadrant and take alternate 32-column chunks; the load of a warp's next chunk is in flight while it stores
        // the current one, and the accumulator is handed back to the MMA warp as soon as its last load has landed.
        const int ew = warp - 4;
        const int q = ew & 3, half = ew >> 2;   // TMEM lanes [32q, 32q+32); chunks half, half+2, ...
        float* sw = sEpi + ew * 32 * 33;
        int as = 0; uint32_t aph = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
            const int tile = w % num_tiles, split = dual ? 0 : w / num_tiles;
            float* const Cout = (dual && w >= num_tiles) ? p.C2 : p.C;
            float* const ssq = (dual && w >= num_tiles) ? p.sumsq_out2 : p.sumsq_out;
            const int mt_n = (MT == 2 && (tile % p.tiles_m) * TM + GBM < p.M) ? 2 : 1;
            const int n0 = (tile / p.tiles_m) * GBN;
            const bool has_k = split * kb_per < num_kb;
            const bool add_bias = p.bias != nullptr && split == 0;
            const int nchunks = mt_n * (GBN / 32);
            float ss = 0.f;
            auto load_chunk = [&](int cc, uint32_t (&v)[32]) {
                const int mt = cc / (GBN / 32), c = cc % (GBN / 32);
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (as * MT + mt) * GBN + c * 32, v);
            };
            auto release_acc = [&]() {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
            };
            auto store_chunk = [&](int cc, const uint32_t (&v)[32]) {
                const int mt = cc / (GBN / 32), c = cc % (GBN / 32);
                const int m0 = (tile % p.tiles_m) * TM + mt * GBM;
                const int nb = n0 + c * 32;
                if (nb >= p.N || m0 + q * 32 >= p.M) return;            // warp-uniform
#pragma unroll
                for (int j = 0; j < 32; ++j) sw[lane * 33 + j] = has_k ? p.alpha * __uint_as_float(v[j]) : 0.f;
                __syncwarp();
                const int col = nb + lane;
                const float bv = (add_bias && col < p.N) ? p.bias[col] + (p.bias2 ? p.bias2[col] : 0.f) : 0.f;
                const int rows = min(32, p.M - (m0 + q * 32));
                if (col < p.N) {
                    float* cptr = Cout + (int64_t)(m0 + q * 32) * p.ldc + col;
                    if (p.splits > 1) {
                        if (rows == 32) {
#pragma unroll
                            for (int r = 0; r < 32; ++r) atomicAdd(cptr + (int64_t)r * p.ldc, sw[r * 33 + lane] + bv);
                        } else {
                            for (int r = 0; r < rows; ++r) atomicAdd(cptr + (int64_t)r * p.ldc, sw[r * 33 + lane] + bv);
                        }
                    } else if (p.accumulate) {
                        for (int r = 0; r < rows; ++r) cptr[(int64_t)r * p.ldc] += sw[r * 33 + lane] + bv;
                    } else if (rows == 32) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const float o = sw[r * 33 + lane] + bv;
                            cptr[(int64_t)r * p.ldc] = o;
                            ss += o * o;
                        }
                    } else {
                        for (int r = 0; r < rows; ++r) {
                            const float o = sw[r * 33 + lane] + bv;
                            cptr[(int64_t)r * p.ldc] = o;
                            ss += o * o;
                        }
                    }
                }
                __syncwarp();
            };
            mbar_wait(&acc_full[as], aph);
            tcgen05_fence_after();
            uint32_t va[32], vb[32];
            int cc = half;                       // nchunks >= 4: every warp owns at least two chunks
            load_chunk(cc, va);
            while (true) {
                tmem_ld_wait();
                const bool more_b = cc + 2 < nchunks;
                if (more_b) load_chunk(cc + 2, vb); else release_acc();
                store_chunk(cc, va);
                if (!more_b) break;
                tmem_ld_wait();
                const bool more_a = cc + 4 < nchunks;
                if (more_a) load_chunk(cc + 4, va); else release_acc();
                store_chunk(cc + 2, vb);
                if (!more_a) break;
                cc += 4;
            }
            if (ssq) {
                ss = warp_sum(ss);
                if (lane == 0) ssq[tile * 8 + ew] = ss;
            }
            if (++as == kAccStages) { as = 0; aph ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    // ONE CTA keeps the grid from completing before the primary has: a CTA blocked here holds its SM, and if every
    // CTA waited, the ~20 SMs the recurrence leaves idle would each run a single tile and then sit until it ends
    // (measured: tools/micro/pdl_overlap.cu, 20 of 148 dependent CTAs ran during the primary)
    if (p.pdl_tail && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- host side ----------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int g_num_sms = 0;
bool g_attr_set[64][16] = {};   // per device: function attributes belong to the device's context

}  // namespace

int tc_num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// fp16 2-D tensor map: `inner` contiguous elements per row, `outer` rows, row pitch ld elements
int tc_make_tmap_f16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                     uint32_t box_outer, int swizzle128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return ZRB_E_CUDA;
    }
    ZRB_REQUIRE((ld * 2) % 16 == 0 && (((uintptr_t)ptr) & 15) == 0, "TMA operand needs 16-byte aligned base and pitch");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return ZRB_E_CUDA;
    }
    return ZRB_OK;
}

template <bool A_MN, bool B_MN, int GBN, int MT>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const GemmArgs& a,
                       cudaStream_t s) {
    auto kern = gemm_f16_tc_kernel<A_MN, B_MN, GBN, MT>;
    const int idx = (A_MN ? 2 : 0) + (B_MN ? 1 : 0) + (GBN == 256 ? 4 : 0) + (MT == 2 ? 8 : 0);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!g_attr_set[dev][idx]) {
        ZRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<GBN, MT>::kSmem));
        g_attr_set[dev][idx] = true;
    }
    int grid = a.tiles_m * a.tiles_n * (a.C2 ? 2 : a.splits);
    if (grid > tc_num_sms()) grid = tc_num_sms();
    if (a.pdl_tail) {
        // programmatic dependent launch: the grid may start as soon as every CTA of the preceding kernel has executed
        // griddepcontrol.launch_dependents (the persistent recurrence kernels do so once they are all resident), and
        // fills the SMs that kernel leaves idle; CTAs that find no free SM start when it ends.  ONE work item per CTA
        // here (not the persistent one-CTA-per-SM grid): the hardware block scheduler then hands tiles to whichever
        // SMs are free, so the ~20 idle SMs work through most of the tiles while the recurrence runs, instead of each
        // late CTA still owning a full static share of them.
        grid = a.tiles_m * a.tiles_n * (a.C2 ? 2 : a.splits);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = GemmCfg<GBN, MT>::kSmem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ZRB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tb2, a));
        count_launch();
        return ZRB_OK;
    }
    kern<<<grid, kGemmThreads, GemmCfg<GBN, MT>::kSmem, s>>>(ta, tb, tb2, a);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

template <int GBN, int MT>
static int dispatch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const GemmArgs& a,
                         int a_mn, int b_mn, cudaStream_t s) {
    if (!a_mn && !b_mn) return launch_gemm<false, false, GBN, MT>(ta, tb, tb2, a, s);
    if (!a_mn && b_mn) return launch_gemm<false, true, GBN, MT>(ta, tb, tb2, a, s);
    if (a_mn && !b_mn) return launch_gemm<true, false, GBN, MT>(ta, tb, tb2, a, s);
    return launch_gemm<true, true, GBN, MT>(ta, tb, tb2, a, s);
}

// Tile shape for an [M,N] output with K-block count num_kb: 128x256 when those tiles give every SM (nearly) a
// full wave, else 128x128; few output tiles but a long contraction (the dgrads) split K in two.
// ZRB_GEMM_MT=2 opts into 256x256 tiles (mt = 2, contraction split up to 8 ways when there are few tiles).  They
// move a third fewer bytes per flop, but measured SLOWER on every shape of this path
// (profiles/r01_gemm_tile_256x256_vs_128x256.json: e.g. [700x10000x1500] 37.2 vs 35.0 us, [6000x1500x700] 29.4 vs
// 27.2 us, [700x6000x1500] 36.3 vs 22.7 us): each SM sustains only ~30 B/clk of 2-D TMA loads whatever the tile, so
// what counts is how many SMs pull at once, and the single accumulator stage exposes the epilogue.
struct TileChoice { int mt, bn, tiles_m, tiles_n, splits; };
static TileChoice choose_tiles(int M, int N, int num_kb, bool can_split) {
    static const bool mt2 = [] { const char* e = getenv("ZRB_GEMM_MT"); return e && e[0] == '2'; }();
    const int nsm = tc_num_sms();
    TileChoice c;
    const int t2 = cdiv(M, 2 * GBM) * cdiv(N, 256);
    int sp2 = 1;
    if (can_split) while (sp2 < 8 && t2 * (sp2 * 2) <= nsm && num_kb / (sp2 * 2) >= 8) sp2 *= 2;
    if (mt2 && M > GBM && t2 * sp2 >= (nsm * 45) / 100) {
        c.mt = 2; c.bn = 256; c.tiles_m = cdiv(M, 2 * GBM); c.tiles_n = cdiv(N, 256); c.splits = sp2;
        return c;
    }
    c.mt = 1;
    c.tiles_m = cdiv(M, GBM);
    const int t256 = c.tiles_m * cdiv(N, 256);
    if (can_split && t256 * 4 <= nsm && t256 * 4 >= (nsm * 8) / 10 && num_kb / 4 >= 8) {
        // the dgrads ([700 x 1500] outputs, 94..157 K blocks): 36 tiles of 128x256 with the contraction split in
        // four move 25% fewer operand bytes per CTA than 72 tiles of 128x128 split in two
        c.bn = 256; c.tiles_n = cdiv(N, 256); c.splits = 4;
        return c;
    }
    c.bn = (t256 >= (nsm * 9) / 10) ? 256 : 128;
    static const int force_bn = [] { const char* e = getenv("ZRB_GEMM_BN"); return e ? atoi(e) : 0; }();   // experiment switch
    if (force_bn == 128 || force_bn == 256) c.bn = force_bn;
    c.tiles_n = cdiv(N, c.bn);
    // few output tiles but a long contraction: split K in two so that ~all SMs work; partials added into a
    // zeroed C with atomics
    c.splits = (can_split && c.tiles_m * c.tiles_n * 2 <= nsm && num_kb >= 8) ? 2 : 1;
    return c;
}

// number of sumsq_out slots gemm_f16_tc writes for an [M,N] output with contraction length K
int gemm_f16_tc_sumsq_slots(int M, int N, int K) {
    TileChoice c = choose_tiles(M, N, cdiv(K, GBK), false);
    return c.tiles_m * c.tiles_n * kEpiWarps;
}

int gemm_f16_tc(const __half* A, int64_t lda, int a_mn, const __half* B, int64_t ldb, int b_mn, float* C, int64_t ldc,
                int M, int N, int K, float alpha, const float* bias, int accumulate, cudaStream_t s, float* sumsq_out,
                const float* bias2, bool pdl, const __half* B2, float* C2, float* sumsq_out2, const __half* A_tiled,
                int a_nt128, const __half* B_tiled, int b_nt128) {
    if (M <= 0 || N <= 0) return ZRB_OK;
    ZRB_REQUIRE(!B2 == !C2, "dual launch needs both B2 and C2");
    ZRB_REQUIRE(!bias2 || bias, "bias2 needs bias");
    ZRB_REQUIRE(!sumsq_out || !accumulate, "sumsq_out needs a plain store epilogue");
    ZRB_REQUIRE(K > 0, "gemm_f16_tc needs K > 0");
    const TileChoice tc = choose_tiles(M, N, cdiv(K, GBK), !sumsq_out && ldc == N && !C2);
    const int bn = tc.bn;
    CUtensorMap ta, tb;
    if (!a_mn) ZRB_TRY(tc_make_tmap_f16(&ta, A, K, M, lda, GBK, GBM, 1));
    else       ZRB_TRY(tc_make_tmap_f16(&ta, A, M, K, lda, 64, GBK, 1));
    if (!b_mn) ZRB_TRY(tc_make_tmap_f16(&tb, B, K, N, ldb, GBK, bn, 1));
    else       ZRB_TRY(tc_make_tmap_f16(&tb, B, N, K, ldb, 64, GBK, 1));
    CUtensorMap tb2 = tb;
    if (B2) {
        if (!b_mn) ZRB_TRY(tc_make_tmap_f16(&tb2, B2, K, N, ldb, GBK, bn, 1));
        else       ZRB_TRY(tc_make_tmap_f16(&tb2, B2, N, K, ldb, 64, GBK, 1));
    }
    GemmArgs a;
    a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.bias = bias; a.C = C; a.ldc = ldc; a.accumulate = accumulate;
    a.tiles_m = tc.tiles_m; a.tiles_n = tc.tiles_n;
    a.splits = tc.splits;
    a.sumsq_out = sumsq_out;
    a.bias2 = bias2;
    a.C2 = C2; a.sumsq_out2 = sumsq_out2;
    a.a_tiled = a_mn ? nullptr : A_tiled; a.a_nt128 = a_nt128; a.b_tiled = b_mn ? nullptr : B_tiled; a.b_nt128 = b_nt128;
    a.pdl_tail = (pdl && a.splits == 1) ? 1 : 0;    // (a split launch is preceded by a memset: nothing to chain to)
    a.pdl_trigger = (rec_pdl_enabled() && !a.pdl_tail) ? 1 : 0;
    // split partials are added into a zeroed C: order-independent for two (a+b == b+a), last-bit run-to-run
    // differences beyond that
    if (a.splits > 1 && !accumulate) ZRB_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s));
    if (tc.mt == 2) return dispatch_gemm<256, 2>(ta, tb, tb2, a, a_mn, b_mn, s);
    return bn == 256 ? dispatch_gemm<256, 1>(ta, tb, tb2, a, a_mn, b_mn, s)
                     : dispatch_gemm<128, 1>(ta, tb, tb2, a, a_mn, b_mn, s);
}

}  // namespace zrb

// EXPERIMENT: the same GEMM (K-major A and B) with both operands given as pre-tiled images as well; a_nt128 / b_nt128 =
// 128-row tiles per K block in the images (b_nt128 even)
extern "C" int zrb_gemm_f16_tiled(const void* A, int64_t lda, const void* B, int64_t ldb, const void* A_tiled, int32_t a_nt128,
                                  const void* B_tiled, int32_t b_nt128, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                                  float alpha, void* stream) {
    ZRB_REQUIRE(A && B && C, "null argument");
    return zrb::gemm_f16_tc((const __half*)A, lda, 0, (const __half*)B, ldb, 0, C, ldc, M, N, K, alpha, nullptr, 0,
                            (cudaStream_t)stream, nullptr, nullptr, false, nullptr, nullptr, nullptr, (const __half*)A_tiled,
                            a_nt128, (const __half*)B_tiled, b_nt128);
}

extern "C" int zrb_gemm_f16(const void* A, int64_t lda, int32_t a_mn_major, const void* B, int64_t ldb,
                            int32_t b_mn_major, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, float alpha,
                            const float* bias, int32_t accumulate, void* stream) {
    ZRB_REQUIRE(A && B && C, "null argument");
    return zrb::gemm_f16_tc((const __half*)A, lda, a_mn_major, (const __half*)B, ldb, b_mn_major, C, ldc, M, N, K,
                            alpha, bias, accumulate, (cudaStream_t)stream, nullptr, nullptr, false, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0);
}

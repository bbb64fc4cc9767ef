// Persistent LSTM recurrence, forward, one launch per layer (sm_100a, cooperative launch).
//
//   for t in 0..T-1:   gates_t = XG_t + h_{t-1} * W_hh^T ;  (i,f,g,o) ;  c_t, h_t     (model.py:34-45)
//
// Work split: CTA k owns U hidden units j in [k*U, k*U+U) (their cell math, c_t in registers for the whole
// window).  The slice of W_hh a CTA multiplies with (fp16, ~144 KB) is loaded ONCE into shared memory in the
// canonical no-swizzle K-major UMMA layout and stays there for all T steps (weight-stationary):
//   SPLIT = false  the 4U gate rows of its own units x the whole contraction, M = 64 tiles   (H < 256)
//   SPLIT = true   CTA PAIRS (clusters of 2): the 8U gate rows of the pair's units x ONE HALF of the contraction,
//                  M = 128 tiles; see the note above RecFwdArgs.  The description below is the SPLIT = false flow;
//                  with SPLIT the drain pushes rows to their owner instead of staging them.
//
// Per step:
//   loader thread   polls the grid-barrier counter (relaxed loads, one fence.acquire.gpu after the last
//                   arrival, fence.proxy.async.global), then brings the 72 KB h_{t-1} operand image (written
//                   by all CTAs, already in UMMA layout; step 0: the image fwd_prep built from the incoming
//                   state) into shared memory as four cp.async.bulk pieces, each with its own mbarrier, so
//                   the MMAs start when the first quarter of K has landed
//   4 MMA threads   H/16 tcgen05.mma (M=64, N=pad8(B), K=16); issuer i takes K steps i, i+4, ... into its
//                   own TMEM accumulator.  One thread issuing back to back pays >= 44.6 clk per instruction
//                   for any N <= 64 (profiles/r01_tcgen05_mma_issue_microbench.csv) and ~90 clk in a real
//                   loop; four issuers reach 33 clk per MMA, the rate at which the tensor core fetches the
//                   2.75 KB of operands of such an instruction from shared memory
//   8 epilogue warps drain the accumulators (each warp sums all four of its lane quadrant / column group and
//                   stages one value per row), add the x-part pre-activations prefetched during the MMAs,
//                   apply sigmoid/tanh/cell update/dropout.  The next step's operand image is stored FIRST
//                   and published (one red.release.gpu on the grid-barrier counter, which is never reset:
//                   the launch gets its starting value); everything backward needs (activated gates, c_t,
//                   row-major fp16 h, dropout(h) for the next layer) is stored after the arrival, off the
//                   critical path.
//
// Roofline: latency/L2/shared-memory bound, not tensor bound -- per step each CTA streams its 144 KB
// weight slice from shared memory through the tensor core (>= 1150 clk at 128 B/clk) and all CTAs
// re-read the 72 KB h image from L2; algorithmic flops per layer call = 8*T*B*H^2.
#include <stdlib.h>

#include "rec_common.cuh"

namespace zrb {

// K-split variant (RecPlan::KS == 2, used when the shape allows): the MMA phase is 1 instruction per K step of 16
// whatever the tile height (M = 64 and M = 128 cost the same ~33 clk at N <= 48), so a CTA that owns 4U = 48 gate rows and
// the whole contraction issues H/16 = 94 half-empty M=64 instructions per step.  A CLUSTER OF TWO CTAs instead owns 2U
// units = 96 gate rows (M = 128, N = 32): CTA r keeps the K half r of all 96 rows resident (same 144 KB), loads only
// its half of the h image, issues 47 instructions, and pushes each accumulator row straight from registers into the
// shared memory of the CTA that owns the row's unit (st.async, bytes counted on the owner's mbarrier: no fence, no
// staging pass); the owner adds the two partial sums in its cell math.
struct RecFwdArgs {
    const __half* w_img;      // [nCTA][KcS][G][8][8]  (K-split: CTA = (pair, K half))
    const __half* h0_img;     // [Kc][GB][8][8] image of the state entering the window: the B operand of step 0
    __half* h_img;            // [T+1][Kc][GB][8][8]; image t (t >= 1) is the B operand of step t, written by step t-1
    float* gates;             // [N,4H] in: x-part pre-activations (+biases); out: activated gates
    const float* c0;          // [B,H]
    float* cst;               // [N,H]
    float* h_last;            // [B,H] or null
    float* c_last;            // [B,H] or null
    __half* hprev_h;          // [N+B,Hp] row-major, rows B.. written here
    __half* y_h;              // [N,Hp] row-major dropout(h)
    float* h_f32;             // or null: [N,H] fp32 h_t (the unit-level entry point zrb_lstm_layer_fwd returns it)
    unsigned int* counter;    // grid barrier: never reset, `base` is its value when this launch starts
    unsigned int base;
    int T, B, H, Hp, U, G, GB, Kc, nCTA;
    int KcS, GBi;             // K chunks per CTA (Kc / KS); 8-row batch groups of the operand image (GB, or 4 when N = 32)
    MaskSrc m;
    RecWatch w;               // watchdog (rec_common.cuh)
    long long* trace;         // optional (profiling): [8] launch stamps (rec_launch_stamps) + [T][8] clock64 stamps of CTA 0
};

__device__ __forceinline__ uint32_t fwd_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t fwd_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void fwd_st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void fwd_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool SPLIT>
__global__ void __launch_bounds__(kRecThreads, 1) lstm_rec_fwd_kernel(RecFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int a_bytes = a.KcS * a.G * 128;     // this CTA's weight slice
    const int b_bytes = a.KcS * a.GBi * 128;   // the part of the h image this CTA multiplies with
    const int Bp = a.GBi * 8;                  // N of the MMA
    const int ldd = Bp + 1;
    const int ldr = Bp + 4;                    // K-split: pitch of the receive buffer rows (16-byte aligned)
    uint8_t* sA = smem;
    uint8_t* sB = smem + a_bytes;
    float* sD = (float*)(sB + b_bytes);        // [64][Bp+1] accumulator staging (sized for two) / K-split: receive buffer
    float* sR = sD;                            //   sR[source rank][gate * U + unit][batch]
    uint64_t* bars = (uint64_t*)((uint8_t*)sD + 2 * 64 * ldd * 4);
    uint64_t* bar_a = bars;        // weight slice landed
    uint64_t* bar_b = bars + 1;    // [kRecPieces] h image pieces of this step landed
    uint64_t* bar_mma = bars + 1 + kRecPieces;  // accumulators ready
    uint64_t* bar_recv = bar_mma + 1;           // K-split: both CTAs' partial sums of my rows have landed
    uint32_t* tmem_slot = (uint32_t*)(bar_recv + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const uint32_t rank = SPLIT ? fwd_cluster_ctarank() : 0u;   // K half this CTA multiplies; also which units it owns
    const int j0 = cta * a.U;                  // (K-split: cta = 2 * pair + rank, the pair owns units [pair*2U, pair*2U + 2U))
    const int nu = max(0, min(a.U, a.H - j0));
    const int ksteps = a.KcS / 2;
    const int piece_steps = (ksteps + kRecPieces - 1) / kRecPieces;
    const bool tr = a.trace != nullptr && cta == 0;
    long long* const trs = a.trace + 8;
    if (a.trace && threadIdx.x == 0) rec_launch_stamps(a.trace, tr, false);

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (int i = 0; i < kRecPieces; ++i) mbar_init(&bar_b[i], 1);
        mbar_init(bar_mma, kRecMmaWarps);
        mbar_init(bar_recv, 1);
        fence_mbar_init();
    }
    if (warp == kRecMmaWarp) tmem_alloc<kRecTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    if (SPLIT) fwd_cluster_sync();   // the partner's mbarriers are initialised before any st.async targets them

    if (warp == kRecLoadWarp && lane == 0) {
        // ===================== loader =====================
        const uint8_t* src = (const uint8_t*)a.w_img + (size_t)cta * a_bytes;
        mbar_expect_tx(bar_a, a_bytes);
        for (int off = 0; off < a_bytes; off += 32768) bulk_load_1d(sA + off, src + off, min(32768, a_bytes - off), bar_a);
        pdl_wait();   // everything below reads what the preceding kernel wrote
        bool dead = false;
        const int lbo_b = a.GBi * 128;
        const size_t img_bytes = (size_t)a.Kc * a.GBi * 128;   // one whole h image; this CTA reads K chunks [rank*KcS, +KcS)
        for (int t = 0; t < a.T; ++t) {
            if (t > 0) grid_counter_wait(a.counter, a.base + (unsigned int)t * a.nCTA, a.w, dead, t);
            if (dead) break;   // (watchdog: a thread that gave up starts no further asynchronous operation)
            if (tr) trs[t * 8 + 0] = clock64();
            fence_proxy_async_global();
            const uint8_t* img = (t == 0 ? (const uint8_t*)a.h0_img : (const uint8_t*)a.h_img + (size_t)t * img_bytes) +
                                 (size_t)rank * b_bytes;
            for (int pc = 0; pc < kRecPieces; ++pc) {
                const int k0 = pc * piece_steps, k1 = min(ksteps, k0 + piece_steps);
                if (k0 >= k1) { mbar_arrive(&bar_b[pc]); continue; }
                const int off = k0 * 2 * lbo_b, bytes = (k1 - k0) * 2 * lbo_b;
                mbar_expect_tx(&bar_b[pc], bytes);
                bulk_load_1d(sB + off, img + off, bytes, &bar_b[pc]);
            }
        }
    } else if (warp >= kRecMmaWarp && warp < kRecMmaWarp + kRecMmaWarps && lane == 0) {
        // ===================== MMA issuers: issuer i takes K steps i, i+2, ... into accumulator i =====================
        const int me = warp - kRecMmaWarp;
        const uint32_t my_acc = tmem_d + me * 32;
        const uint32_t idesc = make_idesc_f16(SPLIT ? 128 : 64, Bp, 0, 0);
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
        const uint32_t lbo_a = a.G * 128, lbo_b = a.GBi * 128;
        bool dead = false;
        bounded_mbar_wait(bar_a, 0, a.w, dead, kWaitWeights, 0);
        for (int t = 0; t < a.T && !dead; ++t) {
            for (int pc = 0; pc < kRecPieces; ++pc) {
                bounded_mbar_wait(&bar_b[pc], t & 1, a.w, dead, kWaitOperand, t);
                if (dead) break;
                tcgen05_fence_after();
                if (tr && pc == 0 && me == 0) trs[t * 8 + 1] = clock64();
                const int k0 = pc * piece_steps, k1 = min(ksteps, k0 + piece_steps);
                for (int ks = k0 + ((me - k0) & (kRecMmaWarps - 1)); ks < k1; ks += kRecMmaWarps) {
                    uint64_t da = make_smem_desc(a_addr + ks * 2 * lbo_a, lbo_a, 128, kSwizzleNone);
                    uint64_t db = make_smem_desc(b_addr + ks * 2 * lbo_b, lbo_b, 128, kSwizzleNone);
                    umma_f16(my_acc, da, db, idesc, ks >= kRecMmaWarps ? 1u : 0u);
                }
            }
            if (!dead) umma_commit(bar_mma);
            if (tr && me == 0) trs[t * 8 + 2] = clock64();
        }
    } else if (warp < kRecEpiWarps) {
        pdl_wait();
        if (threadIdx.x == 0) pdl_launch_dependents();   // after the wait: dependents of this kernel keep stream order with its predecessor
        // ===================== epilogue: 256 threads =====================
        const int tid = threadIdx.x;
        const int B = a.B, H = a.H;
        bool dead = false;
        const int cells = a.U * B;                     // cell = b * U + u (u fastest: contiguous j)
        const uint64_t n_total = (uint64_t)a.T * B * H;
        float creg[kRecMaxCell];
#pragma unroll
        for (int k = 0; k < kRecMaxCell; ++k) {
            int cell = tid + kRecEpiThreads * k;
            int b = cell / a.U, u = cell % a.U;
            creg[k] = (cell < cells && u < nu) ? a.c0[(size_t)b * H + j0 + u] : 0.f;
        }
        const uint32_t sR_addr = smem_u32(sR), bar_recv_addr = smem_u32(bar_recv);
        const int rows_pair = 8 * a.U;                                   // K-split: gate rows of the pair (4 x 2U)
        const uint32_t recv_bytes = 2u * 4u * (uint32_t)a.U * (uint32_t)Bp * 4u;   // 2 sources x 4U rows x Bp columns
        for (int t = 0; t < a.T; ++t) {
            if (SPLIT && tid == 0 && !dead) mbar_expect_tx(bar_recv, recv_bytes);
            // prefetch the x-part pre-activations of this step while the MMAs run
            float pre[kRecMaxCell][4];
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    pre[k][q] = ok ? __ldg(a.gates + ((size_t)t * B + b) * 4 * H + (size_t)q * H + j0 + u) : 0.f;
            }
            bounded_mbar_wait(bar_mma, t & 1, a.w, dead, kWaitAcc, t);
            tcgen05_fence_after();
            if (tr && tid == 0) trs[t * 8 + 3] = clock64();
            {   // 8 warps share the (TMEM lane quadrant, 8-column group) tasks; each sums ALL issuers' accumulators
                // (an issuer with no K step leaves its accumulator unwritten: skipped by a warp-uniform test).
                // Accumulator row i sits in lane (i % 16) + 32 * (i / 16).
                // (A warp can only read the TMEM lane quadrant (warp % 4): task -> quadrant is fixed by the warp index, so the
                // drain order ACROSS quadrants cannot be chosen -- an attempt to send the partner's rows first broke this.)
                for (int task = warp; task < 4 * a.GBi; task += kRecEpiWarps) {
                    const int quad = task & 3, c0 = (task >> 2) * 8;
                    if (SPLIT && 32 * quad >= rows_pair) continue;        // M = 128: row i sits in lane i; padding quadrant
                    uint32_t v[kRecMmaWarps][8];
                    const uint32_t base = tmem_d + ((uint32_t)(32 * quad) << 16) + c0;
#pragma unroll
                    for (int ai = 0; ai < kRecMmaWarps; ++ai)
                        if (ai < ksteps) tmem_ld_32x8(base + ai * 32, v[ai]);
                    tmem_ld_wait();
                    float acc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
                    for (int ai = 0; ai < kRecMmaWarps; ++ai)
                        if (ai < ksteps) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(v[ai][i]);
                        }
                    if (!SPLIT) {
                        if (lane < 16) {
                            float* dst = sD + (16 * quad + lane) * ldd + c0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) dst[i] = acc[i];
                        }
                    } else {
                        // row = 4 * (unit within the pair) + gate: straight into the shared memory of the owning CTA
                        const int row = 32 * quad + lane;
                        if (row < rows_pair && !dead) {
                            // receive rows are gate-major (q * U + u): the cell threads of a warp (consecutive u) then read
                            // addresses ldr floats apart, 4 banks apart, instead of 4 * ldr (2 distinct banks: 16-way conflicts)
                            const int up = row >> 2, owner = up / a.U, lrow = (row & 3) * a.U + (up - owner * a.U);
                            const uint32_t dst = fwd_mapa(sR_addr + (uint32_t)((((int)rank * 4 * a.U + lrow) * ldr + c0) * 4), owner);
                            const uint32_t rbar = fwd_mapa(bar_recv_addr, owner);
                            fwd_st_async_v4(dst, acc[0], acc[1], acc[2], acc[3], rbar);
                            fwd_st_async_v4(dst + 16, acc[4], acc[5], acc[6], acc[7], rbar);
                        }
                    }
                }
            }
            tcgen05_fence_before();
            if (!SPLIT) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tr && tid == 0) trs[t * 8 + 4] = clock64();
            } else {
                if (tr && tid == 0) trs[t * 8 + 4] = clock64();
                bounded_mbar_wait(bar_recv, t & 1, a.w, dead, kWaitRecv, t);   // both K halves of my 4U rows have landed
            }
            float o_i[kRecMaxCell], o_f[kRecMaxCell], o_g[kRecMaxCell], o_o[kRecMaxCell], o_h[kRecMaxCell];
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
                o_i[k] = o_f[k] = o_g[k] = o_o[k] = o_h[k] = 0.f;
                if (!ok) continue;
                float zi, zf, zg, zo;
                if (!SPLIT) {
                    const float* d0 = sD + (4 * u) * ldd + b;
                    zi = pre[k][0] + d0[0];
                    zf = pre[k][1] + d0[ldd];
                    zg = pre[k][2] + d0[2 * ldd];
                    zo = pre[k][3] + d0[3 * ldd];
                } else {
                    const float* r0 = sR + u * ldr + b;                  // K half 0, gate 0
                    const float* r1 = r0 + 4 * a.U * ldr;                // K half 1
                    const int gs = a.U * ldr;                            // gate stride
                    zi = pre[k][0] + (r0[0] + r1[0]);
                    zf = pre[k][1] + (r0[gs] + r1[gs]);
                    zg = pre[k][2] + (r0[2 * gs] + r1[2 * gs]);
                    zo = pre[k][3] + (r0[3 * gs] + r1[3 * gs]);
                }
                float gi = fast_sigmoid(zi), gf = fast_sigmoid(zf), gg = fast_tanh(zg), go = fast_sigmoid(zo);
                float c = gf * creg[k] + gi * gg;
                float h = go * fast_tanh(c);
                creg[k] = c;
                o_i[k] = gi; o_f[k] = gf; o_g[k] = gg; o_o[k] = go; o_h[k] = h;
                // critical path: the next step's operand image [kc][g][r][e], kc = j/8, e = j%8, g = b/8, r = b%8
                const int j = j0 + u;
                __half* img = a.h_img + (size_t)(t + 1) * ((size_t)a.Kc * a.GBi * 64);
                img[((size_t)(j >> 3) * a.GBi + (b >> 3)) * 64 + (b & 7) * 8 + (j & 7)] = __float2half_rn(h);
            }
            if (tr && tid == 0) trs[t * 8 + 5] = clock64();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                if (tr) trs[t * 8 + 6] = clock64();
                grid_counter_arrive(a.counter);
                if (tr) trs[t * 8 + 7] = clock64();
            }
            // off the critical path: what backward and the next layer read after this kernel
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
                if (!ok) continue;
                const int j = j0 + u;
                const size_t n = (size_t)t * B + b;
                float* grow = a.gates + n * 4 * H + j;
                grow[0] = o_i[k]; grow[H] = o_f[k]; grow[2 * (size_t)H] = o_g[k]; grow[3 * (size_t)H] = o_o[k];
                a.cst[n * H + j] = creg[k];
                a.hprev_h[((size_t)B + n) * a.Hp + j] = __float2half_rn(o_h[k]);
                float y = o_h[k] * mask_mul1(a.m, (uint64_t)n * H + j, n_total);
                a.y_h[n * a.Hp + j] = __float2half_rn(y);
                if (a.h_f32) a.h_f32[n * H + j] = o_h[k];
                if (t == a.T - 1) {
                    if (a.h_last) a.h_last[(size_t)b * H + j] = o_h[k];
                    if (a.c_last) a.c_last[(size_t)b * H + j] = creg[k];
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == kRecMmaWarp) tmem_dealloc<kRecTmemCols>(tmem_d);
    if (SPLIT) fwd_cluster_sync();   // nobody leaves while the partner could still address its shared memory
    if (a.trace && threadIdx.x == 0) rec_launch_stamps(a.trace, tr, true);
}

// ---- weight / state image builders ---------------------------------------------------------------
// w_img[cta][kcl][g][r][e] = half(W_hh[q*H + j, k]) with cta = cluster * KS + rank, row i = g*8 + r = 4*uc + q,
// uc = unit within the cluster (UC = KS * U units), j = cluster*UC + uc, k = (rank*KcS + kcl)*8 + e
__global__ void pack_whh_fwd_kernel(const float* __restrict__ W, __half* __restrict__ img, int H, int UC, int G, int KcS,
                                    int KS, int nCTA) {
    const size_t per_cta = (size_t)KcS * G * 64;
    const size_t total = per_cta * nCTA;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int cta = (int)(idx / per_cta);
        size_t r0 = idx % per_cta;
        int e = (int)(r0 & 7), r = (int)((r0 >> 3) & 7);
        int g = (int)((r0 >> 6) % G), kcl = (int)((r0 >> 6) / G);
        int i = g * 8 + r, uc = i >> 2, q = i & 3;
        int cluster = cta / KS, rank = cta % KS;
        int j = cluster * UC + uc, k = (rank * KcS + kcl) * 8 + e;
        float v = 0.f;
        if (uc < UC && j < H && k < H) v = W[((size_t)q * H + j) * H + k];
        img[idx] = __float2half_rn(v);
    }
}

// ---- host ------------------------------------------------------------------------------------------
size_t rec_smem_bytes(int Kc, int G, int GB) {
    return (size_t)Kc * G * 128 + (size_t)Kc * GB * 128 + 2 * 64 * (GB * 8 + 1) * 4 + 128 /*align*/ + 128 /*bars*/;
}

static bool rec_no_coop() {
    // Profilers (Nsight Compute) refuse cooperative + cluster launches; under one (detected through the injection
    // environment it sets up) or with ZRB_NO_COOP=1 cluster kernels are launched without the cooperative attribute,
    // after an occupancy check that the whole grid fits the device
    static const bool v = getenv("ZRB_NO_COOP") != nullptr || getenv("CUDA_INJECTION64_PATH") != nullptr ||
                          getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NVTX_INJECTION64_PATH") != nullptr;
    return v;
}

int rec_fwd_plan(int H, int B, RecPlan* plan) {
    int nsm = tc_num_sms();
    plan->GB = (B + 7) / 8;
    plan->ok = 0;
    plan->KS = 1;
    if (plan->GB * 8 > 32) return ZRB_OK;  // TMEM accumulators / staging sized for N <= 32
    static const bool no_split = getenv("ZRB_REC_NOSPLIT") != nullptr;   // A/B switch
    // K-split pairs (see the kernel header): M = 128 needs N % 16 == 0 -> image batch groups padded to an even count
    if (!no_split && H >= 256) {
        const int Kp = (H + 31) / 32 * 32, Kc = Kp / 8, KcS = Kc / 2, GBi = (plan->GB + 1) / 2 * 2;
        // first choice: at most one (unit, batch) cell per epilogue thread -- a second pass of the cell loop for a handful
        // of cells doubles the critical path of that warp (measured: U = 13, 260 cells, was slower than U = 12)
        for (int pass = 0; pass < 2; ++pass)
            for (int U = 16; U >= 1; --U) {
                const int npair = (H + 2 * U - 1) / (2 * U);
                if (2 * npair > nsm) break;
                if (U * B > (pass == 0 ? 1 : kRecMaxCell) * kRecEpiThreads) continue;
                const int G = U;                              // 8U gate rows of the pair / 8
                const size_t smem = rec_smem_bytes(KcS, G, GBi);
                // M = 128 reads 16 row groups per K chunk: the last chunk reaches (16-G)*128 B past the slice, into the h buffer
                if (smem <= 227 * 1024 && 2 * 4 * U * (GBi * 8 + 4) <= 2 * 64 * (GBi * 8 + 1)) {
                    plan->KS = 2; plan->U = U; plan->G = G; plan->nCTA = 2 * npair; plan->smem = (int)smem;
                    plan->Kc = Kc; plan->KcS = KcS; plan->GBi = GBi; plan->ok = 1;
                    return ZRB_OK;
                }
            }
    }
    int Kp = (H + 15) / 16 * 16;
    plan->Kc = Kp / 8;
    plan->KcS = plan->Kc;
    plan->GBi = plan->GB;
    for (int U = 16; U >= 1; --U) {
        int n = (H + U - 1) / U;
        if (n > nsm) break;
        int G = (4 * U + 7) / 8;
        size_t smem = rec_smem_bytes(plan->Kc, G, plan->GB);
        // the M=64 atom reads 8 row groups per K chunk: the last chunk reaches (8-G)*128 B past the
        // slice, which lands in the h image buffer that follows it
        if (smem <= 227 * 1024 && U * B <= kRecMaxCell * kRecEpiThreads) {
            plan->U = U; plan->G = G; plan->nCTA = n; plan->smem = (int)smem; plan->ok = 1;
            return ZRB_OK;
        }
    }
    return ZRB_OK;
}

int pack_whh_fwd(const float* W, __half* img, int H, const RecPlan& p, cudaStream_t s) {
    pack_whh_fwd_kernel<<<148 * 4, 256, 0, s>>>(W, img, H, p.KS * p.U, p.G, p.KcS, p.KS, p.nCTA);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

int lstm_rec_fwd(const RecPlan& p, const RecWatchdog& wd, const __half* w_img, const __half* h0_img, __half* h_img, float* gates,
                 const float* c0, float* cst, float* h_last, float* c_last, __half* hprev_h, __half* y_h,
                 unsigned int* counter, unsigned int counter_base, int T, int B, int H, int Hp, MaskSrc m, cudaStream_t s,
                 long long* trace, float* h_f32) {
    static bool attr[64] = {};   // per device: function attributes belong to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr[dev]) {
        ZRB_CUDA(cudaFuncSetAttribute(lstm_rec_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZRB_CUDA(cudaFuncSetAttribute(lstm_rec_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    RecFwdArgs a;
    a.w_img = w_img; a.h0_img = h0_img; a.h_img = h_img; a.base = counter_base; a.gates = gates; a.c0 = c0; a.cst = cst; a.h_last = h_last; a.c_last = c_last;
    a.hprev_h = hprev_h; a.y_h = y_h; a.counter = counter; a.h_f32 = h_f32;
    a.T = T; a.B = B; a.H = H; a.Hp = Hp; a.U = p.U; a.G = p.G; a.GB = p.GB; a.Kc = p.Kc; a.nCTA = p.nCTA; a.m = m;
    a.KcS = p.KcS; a.GBi = p.GBi;
    a.trace = trace;
    ZRB_REQUIRE(wd.flag && wd.host, "lstm_rec_fwd needs the context's watchdog words");
    a.w = rec_watch_args(wd);
    a.base += rec_fault_base("fwd");   // (tests only)
    if (trace) ZRB_CUDA(cudaMemsetAsync(trace + 4, 0x80, 2 * sizeof(long long), s));
    if (p.KS == 1) {
        void* args[] = {&a};
        ZRB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_rec_fwd_kernel<false>, dim3(p.nCTA), dim3(kRecThreads), args,
                                             (size_t)p.smem, s));
        count_launch();
        return ZRB_OK;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.nCTA);
    cfg.blockDim = dim3(kRecThreads);
    cfg.dynamicSmemBytes = (size_t)p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cudaError_t e = cudaSuccess;
    // cooperative, or plain + programmatic behind the input GEMM: tc_common.cuh, rec_launch_programmatic()
    const bool programmatic = rec_launch_programmatic(dev) && !trace;
    const bool plain = programmatic || rec_no_coop();
    if (!plain) {
        attrs[1].id = cudaLaunchAttributeCooperative;
        attrs[1].val.cooperative = 1;
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, lstm_rec_fwd_kernel<true>, a);
        if (e == cudaErrorCooperativeLaunchTooLarge) {
            (void)cudaGetLastError();
            set_error("lstm_rec_fwd: the %d-CTA grid cannot be co-resident on this device", p.nCTA);
            return ZRB_E_CUDA;
        }
        if (e != cudaSuccess) (void)cudaGetLastError();
    }
    if (plain || e != cudaSuccess) {
        cfg.numAttrs = 1;
        static int seen_dev = -1, seen_smem = -1, seen_max = 0;   // the query is a host call: once per (device, footprint)
        if (seen_dev != dev || seen_smem != p.smem) {
            int max_clusters = 0;
            cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, lstm_rec_fwd_kernel<true>, &cfg);
            if (oe != cudaSuccess) { (void)cudaGetLastError(); max_clusters = 0; }
            seen_dev = dev; seen_smem = p.smem; seen_max = max_clusters;
        }
        if (seen_max * 2 < p.nCTA) {
            set_error("lstm_rec_fwd: %d CTA pairs needed, the device can hold %d at once", p.nCTA / 2, seen_max);
            return ZRB_E_CUDA;
        }
        if (programmatic) {
            attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        e = cudaLaunchKernelEx(&cfg, lstm_rec_fwd_kernel<true>, a);
    }
    if (e != cudaSuccess) {
        set_error("lstm_rec_fwd launch failed: %s", cudaGetErrorString(e));
        return ZRB_E_CUDA;
    }
    count_launch();
    return ZRB_OK;
}

}  // namespace zrb

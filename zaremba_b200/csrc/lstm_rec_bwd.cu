// Persistent LSTM recurrence, backward, one launch per layer (sm_100a, thread-block clusters).
//
//   for t in T-1..0:  dh_t = mask * dY_t + dG_{t+1} * W_hh ;  cell backward -> dG_t, dc      
//
// The contraction dG_{t+1}[B,4H] * W_hh[4H,H] runs over the 4H gate rows.  A CTA that owned only a few
// hidden units would fill 16 of the 64 rows of the smallest tcgen05 tile, so a CLUSTER owns UC units and
// splits the contraction between its CTAs; every CTA keeps its slice W_hh[rows of its share, units]^T
// (K-major, canonical no-swizzle UMMA layout, ~150 KB) resident in shared memory for all T steps.
//   S = 1  clusters of 4: CTA rank = gate, M = 64 tiles, K = H          (small H; the round-1 shape)
//   S = 2  clusters of 8: CTA rank = 2*gate + K half, M = 128 tiles, K = H/2: half the MMA instructions per
//          step (their cost does not depend on M), half the operand image to fetch
// The partial products D_r[UC x B] are exchanged as a reduce-scatter by PUSHING: while draining TMEM, each
// accumulator row is written straight from registers into the shared memory of the CTA that owns the row's
// unit (st.async, the bytes are counted on the owner's mbarrier: no fence, no staging pass), and the owner adds
// the partials in fixed order in its cell math.  (Round 1 staged them locally, announced them with a
// cluster-scope release arrive -- a second MEMBAR.ALL.GPU per step -- and pulled them through DSMEM: 17 % slower,
// still selectable for S = 1 with ZRB_BWD_PULL=1.)  dc lives in registers for the whole window; the bias
// gradients sum_{t,b} dG are accumulated in registers and reduced over the batch at the end of the kernel.
//
// Per step and CTA: a bulk copy of its part of its gate's dG image (47-72 KB, four pieces), H/(16*S)
// tcgen05.mma, the push exchange, U*B cell updates, one grid-barrier arrival.
// Roofline: latency / L2 bound like the forward kernel; flops per layer call 8*T*B*H^2.
#include <stdlib.h>

#include "rec_common.cuh"

namespace zrb {

struct RecBwdArgs {
    const __half* w_img;      // [nCluster][4][Kc][G][8][8]
    __half* g_img;            // [2][4][Kc][GB][8][8] ring: slot (t & 1) holds kGradScale * dG_t per gate
    const float* dy;          // [N,H] grad wrt the layer's dropout'ed output
    const float* gates;       // [N,4H] activated (i,f,g,o)
    const float* cst;         // [N,H]
    const float* c0;          // [B,H]
    __half* dG_h;             // [N,G4p] row-major, kGradScale * dG
    float* db1;               // [4H] or null: bias gradient sum_{t,b} dG (model.py:35-36: b_ih and b_hh get the same
    float* db2;               //      gradient), accumulated in registers over the window and reduced over the batch here
    float* db_scratch;        // [4][B][H] fp32 scratch of that reduction (needed when db1 is set)
    int push;                 // exchange of the cluster's partial products: 1 = st.async pushes into the owners' shared
                              // memory (complete_tx on their mbarrier), 0 = stage + remote arrive + DSMEM pulls
    unsigned int* res_flag;   // or null: CTA 0 publishes res_value here when the whole grid is resident
    unsigned int res_value;
    unsigned int* counter;    // grid barrier: never reset, `base` is its value when this launch starts
    unsigned int base;
    int T, B, H, G4p, U, G, GB, Kc, nCTA;
    int KcS, GBi;             // K chunks per CTA (Kc / S); 8-row batch groups of the dG images (GB, or 4 when N = 32)
    MaskSrc m;
    RecWatch w;               // watchdog (rec_common.cuh)
    long long* trace;         // optional (profiling): [8] launch stamps (rec_launch_stamps) + [T][8] clock64 stamps of CTA 0
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr));
    return v;
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// 16 bytes into a peer CTA's shared memory; the bytes are counted on that CTA's mbarrier (complete_tx), so the
// reader needs no fence: observing the phase completion makes them visible
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// S = 1: clusters of 4 (CTA rank = gate), M = 64 tiles, the whole gate block as contraction (S = 1 also keeps the
//        staging + DSMEM-pull exchange selectable with a.push = 0).
// S = 2: clusters of 8.  The MMA phase costs one instruction per K step whatever the tile height, so the cluster owns
//        twice the units (8U = 96 rows, M = 128, N = 32) and CTA rank r = 2*gate + half multiplies only HALF of its gate's
//        rows: 47 instructions per step instead of 94, half the operand image to fetch; the eight partial products are
//        pushed (st.async) into the owners' shared memory and summed in fixed order.
template <int S>
__global__ void __launch_bounds__(kRecThreads, 1) lstm_rec_bwd_kernel(RecBwdArgs a) {
    constexpr int CS = 4 * S;   // cluster size
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int a_bytes = a.KcS * a.G * 128;     // this CTA's weight slice
    const int b_bytes = a.KcS * a.GBi * 128;   // the part of its gate's dG image this CTA multiplies with
    const int Bp = a.GBi * 8;                  // N of the MMA
    const int ldd = Bp + 1;
    uint8_t* sA = smem;
    uint8_t* sB = smem + a_bytes;
    float* sD = (float*)(sB + b_bytes);  // [64][Bp+1] this CTA's partial product, all issuers' accumulators summed (sized for two)
    uint64_t* bars = (uint64_t*)((uint8_t*)sD + 2 * 64 * ldd * 4);
    uint64_t* bar_a = bars;
    uint64_t* bar_b = bars + 1;                    // [kRecPieces]
    uint64_t* bar_mma = bars + 1 + kRecPieces;
    uint64_t* bar_part = bar_mma + 1;              // 4 arrivals per step: every CTA of the cluster staged its partial
    uint64_t* bar_recv = bar_part + 1;             // push mode: all four CTAs' partials of this CTA's units have landed
    uint32_t* tmem_slot = (uint32_t*)(bar_recv + 1);
    // push mode reuses the staging buffer as the receive buffer sR[source rank][unit][batch (pitch ldr, 16-byte rows)]
    const int ldr = Bp + 4;
    float* sR = sD;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x / CS;
    const int UC = CS * a.U;
    const int jc0 = cluster * UC;            // first unit of the cluster
    const int j0 = jc0 + (int)rank * a.U;    // first unit whose cell math this CTA owns
    const int nu = max(0, min(a.U, a.H - j0));
    const int gate = (int)rank / S, khalf = (int)rank % S;   // contraction slice: rows [khalf*KcS*8, +KcS*8) of gate block `gate`
    const int T = a.T, B = a.B, H = a.H;
    const int ksteps = a.KcS / 2;
    const int piece_steps = (ksteps + kRecPieces - 1) / kRecPieces;
    const bool tr = a.trace != nullptr && blockIdx.x == 0;
    long long* const trs = a.trace + 8;
    if (a.trace && threadIdx.x == 0) rec_launch_stamps(a.trace, tr, false);

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (int i = 0; i < kRecPieces; ++i) mbar_init(&bar_b[i], 1);
        mbar_init(bar_mma, kRecMmaWarps);
        mbar_init(bar_part, CS);
        mbar_init(bar_recv, 1);
        fence_mbar_init();
    }
    if (warp == kRecMmaWarp) tmem_alloc<kRecTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    cluster_sync_all();   // every CTA's mbarriers are initialised before any remote arrive

    if (warp == kRecLoadWarp && lane == 0) {
        // ===================== loader =====================
        const uint8_t* src = (const uint8_t*)a.w_img + ((size_t)cluster * CS + rank) * a_bytes;
        mbar_expect_tx(bar_a, a_bytes);
        for (int off = 0; off < a_bytes; off += 32768) bulk_load_1d(sA + off, src + off, min(32768, a_bytes - off), bar_a);
        pdl_wait();   // everything below reads what the preceding kernel wrote
        bool dead = false;
        const int lbo_b = a.GBi * 128;
        const size_t gate_bytes = (size_t)a.Kc * a.GBi * 128;   // one gate's whole dG image
        const bool publish = a.res_flag != nullptr && blockIdx.x == 0;
        if (publish && T == 1) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.res_flag), "r"(a.res_value) : "memory");
        for (int s = 1; s < T; ++s) {
            const int t = T - 1 - s;                      // step being computed; needs dG_{t+1}
            grid_counter_wait(a.counter, a.base + (unsigned int)s * a.nCTA, a.w, dead, s);
            if (publish && s == 1)   // every CTA arrived once: the whole grid is resident (or gave up: a stream gated on this must not hang)
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.res_flag), "r"(a.res_value) : "memory");
            if (dead) break;   // (watchdog: a thread that gave up starts no further asynchronous operation)
            if (tr) trs[s * 8 + 0] = clock64();
            fence_proxy_async_global();
            const uint8_t* img = (const uint8_t*)a.g_img + ((size_t)((t + 1) & 1) * 4 + gate) * gate_bytes +
                                 (size_t)khalf * b_bytes;
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
        const uint32_t idesc = make_idesc_f16(S == 2 ? 128 : 64, Bp, 0, 0);
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
        const uint32_t lbo_a = a.G * 128, lbo_b = a.GBi * 128;
        bool dead = false;
        bounded_mbar_wait(bar_a, 0, a.w, dead, kWaitWeights, 0);
        for (int s = 1; s < T && !dead; ++s) {
            for (int pc = 0; pc < kRecPieces; ++pc) {
                bounded_mbar_wait(&bar_b[pc], (s - 1) & 1, a.w, dead, kWaitOperand, s);
                if (dead) break;
                tcgen05_fence_after();
                if (tr && pc == 0 && me == 0) trs[s * 8 + 1] = clock64();
                const int k0 = pc * piece_steps, k1 = min(ksteps, k0 + piece_steps);
                for (int ks = k0 + ((me - k0) & (kRecMmaWarps - 1)); ks < k1; ks += kRecMmaWarps) {
                    uint64_t da = make_smem_desc(a_addr + ks * 2 * lbo_a, lbo_a, 128, kSwizzleNone);
                    uint64_t db = make_smem_desc(b_addr + ks * 2 * lbo_b, lbo_b, 128, kSwizzleNone);
                    umma_f16(my_acc, da, db, idesc, ks >= kRecMmaWarps ? 1u : 0u);
                }
            }
            if (!dead) umma_commit(bar_mma);
            if (tr && me == 0) trs[s * 8 + 2] = clock64();
        }
    } else if (warp < kRecEpiWarps) {
        pdl_wait();
        if (threadIdx.x == 0) pdl_launch_dependents();   // after the wait: dependents of this kernel keep stream order with its predecessor
        // ===================== epilogue: 256 threads, cells (u, b) of this CTA's U units =====================
        const int tid = threadIdx.x;
        bool dead = false;
        const int cells = a.U * B;                     // cell = b * U + u (u fastest: contiguous j)
        float dcreg[kRecMaxCell], bsum[kRecMaxCell][4];
#pragma unroll
        for (int k = 0; k < kRecMaxCell; ++k) {
            dcreg[k] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) bsum[k][q] = 0.f;
        }
        const uint64_t n_total = (uint64_t)T * B * H;
        const uint32_t sD_addr = smem_u32(sD);
        uint32_t part_addr[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) part_addr[rr] = mapa_shared(sD_addr, rr);   // (pull exchange: S == 1 only)
        const uint32_t bar_part_addr = smem_u32(bar_part);
        const uint32_t sR_addr = smem_u32(sR), bar_recv_addr = smem_u32(bar_recv);
        const uint32_t recv_bytes = (uint32_t)CS * (uint32_t)a.U * (uint32_t)Bp * 4u;   // CS sources x U units x Bp columns
        const float inv = 1.f / kGradScale;
        const size_t img_gate = (size_t)a.Kc * a.GBi * 64;
        const bool push = S == 2 || a.push != 0;

        for (int s = 0; s < T; ++s) {
            const int t = T - 1 - s;
            // prefetch this step's saved activations and upstream gradient
            float gi[kRecMaxCell], gf[kRecMaxCell], gg[kRecMaxCell], go[kRecMaxCell], ct[kRecMaxCell], cp[kRecMaxCell],
                dyv[kRecMaxCell];
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
                gi[k] = gf[k] = gg[k] = go[k] = ct[k] = cp[k] = dyv[k] = 0.f;
                if (ok) {
                    const int j = j0 + u;
                    const size_t n = (size_t)t * B + b;
                    const float* grow = a.gates + n * 4 * H + j;
                    gi[k] = __ldg(grow); gf[k] = __ldg(grow + H); gg[k] = __ldg(grow + 2 * (size_t)H);
                    go[k] = __ldg(grow + 3 * (size_t)H);
                    ct[k] = __ldg(a.cst + n * H + j);
                    cp[k] = t > 0 ? __ldg(a.cst + (n - B) * H + j) : __ldg(a.c0 + (size_t)b * H + j);
                    dyv[k] = __ldg(a.dy + n * H + j) * mask_mul1(a.m, (uint64_t)n * H + j, n_total);
                }
            }
            if (s > 0) {
                if (push && tid == 0 && !dead) mbar_expect_tx(bar_recv, recv_bytes);
                bounded_mbar_wait(bar_mma, (s - 1) & 1, a.w, dead, kWaitAcc, s);
                tcgen05_fence_after();
                if (tr && tid == 0) trs[s * 8 + 3] = clock64();
                // TMEM -> own shared staging: accumulator row i (cluster-local unit) in lane (i%16)+32*(i/16)
                {   // 8 warps share the (TMEM lane quadrant, 8-column group) tasks; each sums ALL issuers' accumulators
                    // (an issuer with no K step leaves its accumulator unwritten: skipped by a warp-uniform test).
                    // Accumulator row i sits in lane (i % 16) + 32 * (i / 16).
                    // (task -> lane quadrant is fixed by the warp index: a warp reads only TMEM lanes [32*(warp%4), +32))
                    for (int task = warp; task < 4 * a.GBi; task += kRecEpiWarps) {
                        const int quad = task & 3, c0 = (task >> 2) * 8;
                        if (S == 2 && 32 * quad >= UC) continue;          // M = 128: row i sits in lane i; padding quadrant
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
                        if (S == 2 || lane < 16) {
                            const int row = S == 2 ? 32 * quad + lane : 16 * quad + lane;   // cluster-local unit of this row
                            if (!push) {
                                float* dst = sD + row * ldd + c0;
#pragma unroll
                                for (int i = 0; i < 8; ++i) dst[i] = acc[i];
                            } else if (row < UC && !dead) {
                                // straight from the registers into the shared memory of the CTA that owns this unit
                                const int owner = row / a.U, uo = row - owner * a.U;
                                const uint32_t dst = mapa_shared(sR_addr + (uint32_t)((((int)rank * a.U + uo) * ldr + c0) * 4), owner);
                                const uint32_t rbar = mapa_shared(bar_recv_addr, owner);
                                st_async_v4(dst, acc[0], acc[1], acc[2], acc[3], rbar);
                                st_async_v4(dst + 16, acc[4], acc[5], acc[6], acc[7], rbar);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                if (!push) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (tr && tid == 0) trs[s * 8 + 4] = clock64();
                    if (tid < 4) mbar_arrive_remote_release(mapa_shared(bar_part_addr, tid));
                    {   // wait until all four CTAs of the cluster staged their partials
                        uint32_t n = 0; long long t0 = 0;
                        while (!dead && !mbar_try_wait_acq_cluster(bar_part, (s - 1) & 1)) {
                            if ((++n & 0xFFFu) == 0 && rec_spin_check(a.w, t0, kWaitPart, s)) dead = true;
                        }
                    }
                } else {
                    if (tr && tid == 0) trs[s * 8 + 4] = clock64();
                    bounded_mbar_wait(bar_recv, (s - 1) & 1, a.w, dead, kWaitRecv, s);   // all CS x U x Bp partial sums of my units have landed
                }
            }
            if (tr && tid == 0) trs[s * 8 + 5] = clock64();
            __half hv[kRecMaxCell][4];
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
                if (!ok) continue;
                float dh = dyv[k];
                if (s > 0) {
                    float pp[CS];
                    if (!push) {
                        const uint32_t off = (uint32_t)(((int)rank * a.U + u) * ldd + b) * 4u;
#pragma unroll             // issue all four DSMEM loads before the first use (each is ~200+ clk)
                        for (int rr = 0; rr < 4; ++rr) pp[rr] = ld_dsmem_f32(part_addr[rr] + off);
                    } else {
#pragma unroll
                        for (int rr = 0; rr < CS; ++rr) pp[rr] = sR[(rr * a.U + u) * ldr + b];
                    }
                    float r = (pp[0] + pp[1]) + (pp[2] + pp[3]);
                    if constexpr (S == 2) r += (pp[4] + pp[5]) + (pp[6] + pp[7]);
                    dh += r * inv;
                }
                const float tc = fast_tanh(ct[k]);
                const float d_o = dh * tc;
                const float dcc = dcreg[k] + dh * go[k] * (1.f - tc * tc);
                const float d_i = dcc * gg[k], d_g = dcc * gi[k], d_f = dcc * cp[k];
                dcreg[k] = dcc * gf[k];
                float dg4[4];
                dg4[0] = d_i * gi[k] * (1.f - gi[k]);
                dg4[1] = d_f * gf[k] * (1.f - gf[k]);
                dg4[2] = d_g * (1.f - gg[k] * gg[k]);
                dg4[3] = d_o * go[k] * (1.f - go[k]);
                const int j = j0 + u;
                // critical path: the four gate images the next step multiplies with
                __half* img = a.g_img + (size_t)(t & 1) * 4 * img_gate + ((size_t)(j >> 3) * a.GBi + (b >> 3)) * 64 +
                              (b & 7) * 8 + (j & 7);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = fminf(fmaxf(dg4[q] * kGradScale, -65504.f), 65504.f);
                    hv[k][q] = __float2half_rn(v);
                    img[(size_t)q * img_gate] = hv[k][q];
                    bsum[k][q] += dg4[q];
                }
            }
            if (tr && tid == 0) trs[s * 8 + 6] = clock64();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                grid_counter_arrive(a.counter);
                if (tr) trs[s * 8 + 7] = clock64();
            }
            // off the critical path: row-major image for the batched dgrad / wgrad GEMMs
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                bool ok = cell < cells && u < nu;
                if (!ok) continue;
                __half* hrow = a.dG_h + ((size_t)t * B + b) * a.G4p + j0 + u;
#pragma unroll
                for (int q = 0; q < 4; ++q) hrow[(size_t)q * H] = hv[k][q];
            }
        }
        if (a.db1) {
            // bias gradients: per-cell sums over the window -> a global scratch [4][B][H] (no shared-memory region of
            // a guaranteed size is free: peers may still read the staging buffer) -> fixed-order sum over the batch by
            // one thread per (gate, unit) of this CTA.  bar.sync orders the CTA's own global writes for its readers.
#pragma unroll
            for (int k = 0; k < kRecMaxCell; ++k) {
                int cell = tid + kRecEpiThreads * k;
                int b = cell / a.U, u = cell % a.U;
                if (cell < cells && u < nu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) a.db_scratch[((size_t)q * B + b) * H + j0 + u] = bsum[k][q];
                }
            }
            __threadfence_block();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < 4 * a.U) {
                const int q = tid / a.U, u = tid % a.U;
                if (u < nu) {
                    float sacc = 0.f;
                    for (int b = 0; b < B; ++b) sacc += a.db_scratch[((size_t)q * B + b) * H + j0 + u];
                    a.db1[(size_t)q * H + j0 + u] = sacc;
                    if (a.db2) a.db2[(size_t)q * H + j0 + u] = sacc;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == kRecMmaWarp) tmem_dealloc<kRecTmemCols>(tmem_d);
    cluster_sync_all();   // no CTA leaves while a peer may still read its staged partial
    if (a.trace && threadIdx.x == 0) rec_launch_stamps(a.trace, tr, true);
}

// w_img[cluster][rank][kcl][g][rr][e] = half(W_hh[gate*H + (khalf*KcS + kcl)*8 + e, cluster*UC + g*8 + rr]) with
// rank = gate*S + khalf: one warp per (cluster, rank, kcl) reads 8 rows x UC contiguous floats and writes one contiguous
// 16*UC-byte block.
__global__ void pack_whh_bwd_kernel(const float* __restrict__ W, __half* __restrict__ img, int H, int UC, int G, int KcS,
                                    int S, int nCluster) {
    __shared__ __half tile[8][8 * 128];
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CS = 4 * S;
    const long long total = (long long)nCluster * CS * KcS;
    for (long long w = (long long)blockIdx.x * 8 + warp_in_block; w < total; w += (long long)gridDim.x * 8) {
        const int kcl = (int)(w % KcS);
        const int r = (int)((w / KcS) % CS);
        const int cl = (int)(w / ((long long)KcS * CS));
        const int gate = r / S, khalf = r % S;
        __half* t = tile[warp_in_block];
        const int rows8 = G * 8;
        for (int idx = lane; idx < 8 * rows8; idx += 32) {
            int e = idx / rows8, u = idx % rows8;            // u fastest: contiguous global reads
            int k = (khalf * KcS + kcl) * 8 + e, j = cl * UC + u;
            float v = (k < H && u < UC && j < H) ? W[((size_t)gate * H + k) * H + j] : 0.f;
            t[u * 8 + e] = __float2half_rn(v);
        }
        __syncwarp();
        __half* dst = img + (((size_t)cl * CS + r) * KcS + kcl) * ((size_t)G * 64);
        for (int idx = lane; idx < rows8 * 8; idx += 32) dst[idx] = t[idx];
        __syncwarp();
    }
}

static bool rec_bwd_no_coop() {
    // Profilers (Nsight Compute) refuse cooperative + cluster launches; under one (detected through the injection
    // environment it sets up) or with ZRB_NO_COOP=1 the kernel is launched as a plain cluster launch after an occupancy
    // check that the whole grid fits the device.  Without the cooperative guarantee another context holding SMs (MPS, a
    // concurrent kernel) could leave CTAs unscheduled; the barrier waits are bounded: after ~3 s the kernel gives up and reports it (rec_common.cuh: RecWatch) instead of hanging.
    static const bool v = getenv("ZRB_NO_COOP") != nullptr || getenv("CUDA_INJECTION64_PATH") != nullptr ||
                          getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NVTX_INJECTION64_PATH") != nullptr;
    return v;
}

template <int S>
static int rec_bwd_max_clusters(int smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * S * 64);
    cfg.blockDim = dim3(kRecThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4 * S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaFuncSetAttribute(lstm_rec_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, lstm_rec_bwd_kernel<S>, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int rec_bwd_plan(int H, int B, RecPlan* plan) {
    int nsm = tc_num_sms();
    plan->GB = (B + 7) / 8;
    plan->ok = 0;
    plan->KS = 1;
    if (plan->GB * 8 > 32) return ZRB_OK;
    static const bool no_split = getenv("ZRB_REC_NOSPLIT") != nullptr;   // A/B switch
    // clusters of 8, half a gate block per CTA (see the kernel header); M = 128 needs N % 16 == 0
    if (!no_split && H >= 256) {
        const int Kp = (H + 31) / 32 * 32, Kc = Kp / 8, KcS = Kc / 2, GBi = (plan->GB + 1) / 2 * 2;
        // first choice: at most one (unit, batch) cell per epilogue thread (see rec_fwd_plan)
        for (int pass = 0; pass < 2; ++pass)
            for (int U = 16; U >= 1; --U) {
                const int UC = 8 * U;
                if (UC > 128) continue;
                const int ncl = (H + UC - 1) / UC;
                if (ncl * 8 > nsm) break;
                if (U * B > (pass == 0 ? 1 : kRecMaxCell) * kRecEpiThreads) continue;
                const int G = UC / 8;
                const size_t smem = rec_smem_bytes(KcS, G, GBi);
                if (smem <= 227 * 1024 && 8 * U * (GBi * 8 + 4) <= 2 * 64 * (GBi * 8 + 1)) {
                    if (rec_bwd_max_clusters<2>((int)smem) < ncl) continue;   // the GPCs cannot hold that many 8-CTA clusters
                    plan->KS = 2; plan->U = U; plan->G = G; plan->nCTA = ncl * 8; plan->smem = (int)smem;
                    plan->Kc = Kc; plan->KcS = KcS; plan->GBi = GBi; plan->ok = 1;
                    return ZRB_OK;
                }
            }
    }
    int Kp = (H + 15) / 16 * 16;
    plan->Kc = Kp / 8;
    plan->KcS = plan->Kc;
    plan->GBi = plan->GB;
    int max_clusters = (nsm - 16) / 4;   // clusters of 4 strand up to 16 SMs (GPC remainders)
    for (int U = 16; U >= 1; --U) {
        int UC = 4 * U;
        if (UC % 8) continue;
        int ncl = (H + UC - 1) / UC;
        if (ncl > max_clusters) break;
        int G = UC / 8;
        size_t smem = rec_smem_bytes(plan->Kc, G, plan->GB);
        if (smem <= 227 * 1024 && U * B <= kRecMaxCell * kRecEpiThreads) {
            plan->U = U; plan->G = G; plan->nCTA = ncl * 4; plan->smem = (int)smem; plan->ok = 1;
            return ZRB_OK;
        }
    }
    return ZRB_OK;
}

int pack_whh_bwd(const float* W, __half* img, int H, const RecPlan& p, cudaStream_t s) {
    const int CS = 4 * p.KS;
    pack_whh_bwd_kernel<<<148 * 4, 256, 0, s>>>(W, img, H, CS * p.U, p.G, p.KcS, p.KS, p.nCTA / CS);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

template <int S>
static int launch_rec_bwd(const RecPlan& p, const RecBwdArgs& a, cudaStream_t s) {
    constexpr int CS = 4 * S;
    static bool attr[64] = {};   // per device: function attributes belong to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr[dev]) {
        ZRB_CUDA(cudaFuncSetAttribute(lstm_rec_bwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.nCTA);
    cfg.blockDim = dim3(kRecThreads);
    cfg.dynamicSmemBytes = (size_t)p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    // The grid barrier needs all nCTA CTAs co-resident.  Either the cooperative attribute makes the driver guarantee it
    // (or refuse the launch), or -- one context on the device, or under a profiler (rec_bwd_no_coop()) -- a plain cluster
    // launch checked against the occupancy query, with the programmatic attribute so that the CTAs start behind the
    // preceding GEMM's trigger (tc_common.cuh, rec_launch_programmatic()).
    const bool programmatic = rec_launch_programmatic(dev) && !a.trace;
    const bool no_coop = programmatic || rec_bwd_no_coop();
    cudaError_t e = cudaSuccess;
    if (!no_coop) {
        attrs[1].id = cudaLaunchAttributeCooperative;
        attrs[1].val.cooperative = 1;
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, lstm_rec_bwd_kernel<S>, a);
        if (e == cudaErrorCooperativeLaunchTooLarge) {
            (void)cudaGetLastError();
            set_error("lstm_rec_bwd: the %d-CTA grid cannot be co-resident on this device (cooperative launch too large)",
                      p.nCTA);
            return ZRB_E_CUDA;
        }
        if (e != cudaSuccess) (void)cudaGetLastError();   // e.g. not supported under a tool: try the checked plain launch
    }
    if (no_coop || e != cudaSuccess) {
        cfg.numAttrs = 1;
        static int seen_dev = -1, seen_smem = -1, seen_max = 0;   // the query is a host call: once per (device, footprint)
        if (seen_dev != dev || seen_smem != p.smem) {
            int max_clusters = 0;
            cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, lstm_rec_bwd_kernel<S>, &cfg);
            if (oe != cudaSuccess) { (void)cudaGetLastError(); max_clusters = 0; }
            seen_dev = dev; seen_smem = p.smem; seen_max = max_clusters;
        }
        if (seen_max * CS < p.nCTA) {
            set_error("lstm_rec_bwd: %d clusters of %d needed, the device can hold %d at once", p.nCTA / CS, CS, seen_max);
            return ZRB_E_CUDA;
        }
        if (programmatic) {
            attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        e = cudaLaunchKernelEx(&cfg, lstm_rec_bwd_kernel<S>, a);
    }
    if (e != cudaSuccess) {
        set_error("lstm_rec_bwd launch failed: %s", cudaGetErrorString(e));
        return ZRB_E_CUDA;
    }
    count_launch();
    return ZRB_OK;
}

int lstm_rec_bwd(const RecPlan& p, const RecWatchdog& wd, const __half* w_img, __half* g_img, const float* dy, const float* gates,
                 const float* cst, const float* c0, __half* dG_h, unsigned int* counter, unsigned int counter_base, int T,
                 int B, int H, int G4p, MaskSrc m, cudaStream_t s, long long* trace, float* db1, float* db2,
                 unsigned int* resident_flag, unsigned int resident_value, float* db_scratch) {
    ZRB_REQUIRE(!db1 || db_scratch, "bias gradients need the scratch buffer");
    RecBwdArgs a;
    a.base = counter_base;
    a.w_img = w_img; a.g_img = g_img; a.dy = dy; a.gates = gates; a.cst = cst; a.c0 = c0; a.dG_h = dG_h;
    static const bool pull = getenv("ZRB_BWD_PULL") != nullptr;   // A/B switch: the r01 staging + DSMEM-pull exchange (S = 1)
    a.push = pull ? 0 : 1;
    a.counter = counter; a.db1 = db1; a.db2 = db2; a.db_scratch = db_scratch; a.res_flag = resident_flag; a.res_value = resident_value;
    a.T = T; a.B = B; a.H = H; a.G4p = G4p; a.U = p.U; a.G = p.G; a.GB = p.GB; a.Kc = p.Kc; a.nCTA = p.nCTA; a.m = m; a.trace = trace;
    a.KcS = p.KcS; a.GBi = p.GBi;
    ZRB_REQUIRE(wd.flag && wd.host, "lstm_rec_bwd needs the context's watchdog words");
    a.w = rec_watch_args(wd);
    a.base += rec_fault_base("bwd");   // (tests only)
    if (trace) ZRB_CUDA(cudaMemsetAsync(trace + 4, 0x80, 2 * sizeof(long long), s));
    return p.KS == 2 ? launch_rec_bwd<2>(p, a, s) : launch_rec_bwd<1>(p, a, s);
}

}  // namespace zrb

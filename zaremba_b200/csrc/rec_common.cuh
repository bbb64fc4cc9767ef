// Helpers shared by the persistent recurrence kernels (forward / backward).
#pragma once
#include <string.h>
#include "tc_common.cuh"
#include "tc_kernels.h"

namespace zrb {
using namespace tc;

constexpr int kRecEpiWarps = 8;                       // warps 0-7: accumulator drain + cell math
constexpr int kRecEpiThreads = kRecEpiWarps * 32;
constexpr int kRecMmaWarp = 8;                        // warps 8..11: lane 0 of warp 8+i issues K steps i, i+4, ... into its
constexpr int kRecMmaWarps = 4;                       //   OWN accumulator i (warp 8 also owns the TMEM allocation).  8 issuers
                                                      //   finish ISSUING in ~2100 clk, but the step did not get shorter:
                                                      //   the MMAs complete at ~33 clk each either way and the drain has
                                                      //   twice the accumulators to read
constexpr int kRecLoadWarp = kRecMmaWarp + kRecMmaWarps;   // lane 0: grid-barrier wait + bulk copies
constexpr int kRecThreads = (kRecLoadWarp + 1) * 32;
constexpr int kRecTmemCols = 32 * kRecMmaWarps;       // one fp32 accumulator (N <= 32 columns) per issuer.  tcgen05.mma has a
                                                      // ~45 clk floor per instruction for N <= 64 (measured); one issuing
                                                      // thread only reaches ~90 clk (descriptor math + R2UR in series with
                                                      // the issue); 2 threads with private accumulators reach ~57, 4 threads ~33 clk per MMA
constexpr int kRecPieces = 4;                         // operand image arrives in this many bulk copies.  (One grid-barrier
                                                      // counter and loader lane PER PIECE, so that a late CTA only delays the
                                                      // piece it writes, was measured 5% slower: 4x the polling traffic and a
                                                      // longer arrival; issuing both drain tasks' TMEM loads before one wait
                                                      // was also slower than two rounds.)
constexpr int kRecMaxCell = 2;                        // (unit, batch) cells per epilogue thread
// ---- watchdog -------------------------------------------------------------------------------------
// Every wait of the persistent kernels is bounded.  A wait that runs out (a lost wake-up, a grid that is not co-resident)
// publishes a code in the context's abort word; from then on EVERY wait of EVERY thread returns at once (a thread that
// has seen the word set stops waiting for good), so the kernel runs to its end through its unchanged barrier skeleton --
// producing garbage, but terminating, with the CUDA context intact.  The host finds the code in a mapped host word at its
// next API call and fails the zrb context (api.cu: watchdog_check).  Nothing on the fast path but a register test.
struct RecWatch {
    unsigned int* flag;       // device word polled by the spinning threads (0 = healthy)
    unsigned int* host;       // mapped host word the first thread to give up writes the code to
    long long spin_cycles;    // ~3 s at 2 GHz unless ZRB_SPIN_CYCLES says otherwise
};
// host: the kernel argument for one launch (ZRB_SPIN_CYCLES shortens the time-out; read once), and the fault injection
// of tests/test_gpu_watchdog.py: ZRB_FAULT_BARRIER_BASE="fwd" / "bwd" makes that kernel's launches expect one arrival
// more than the grid will ever deliver -- every CTA then sits at its first grid barrier like after a lost wake-up.
static inline RecWatch rec_watch_args(const RecWatchdog& wd) {
    static const long long spin = [] { const char* e = getenv("ZRB_SPIN_CYCLES"); long long v = e ? atoll(e) : 0; return v > 0 ? v : 6000000000ll; }();
    RecWatch w;
    w.flag = wd.flag; w.host = wd.host; w.spin_cycles = spin;
    return w;
}
static inline unsigned int rec_fault_base(const char* which) {
    static const char* fault = getenv("ZRB_FAULT_BARRIER_BASE");
    return (fault && !strcmp(fault, which)) ? 1u : 0u;
}
enum { kWaitWeights = 1, kWaitOperand = 2, kWaitAcc = 3, kWaitRecv = 4, kWaitGrid = 5, kWaitPart = 6 };

__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// code = wait kind | CTA << 8 | step << 20
static __device__ __noinline__ void rec_give_up(unsigned int* flag, unsigned int* host, int kind, int step) {
    const unsigned int code = (unsigned int)kind | ((unsigned int)(blockIdx.x & 0xFFF) << 8) | ((unsigned int)(step & 0xFFF) << 20);
    if (atomicCAS(flag, 0u, code) == 0u) {
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(host), "r"(code) : "memory");
        __threadfence_system();
    }
}
// one slow-path visit of a spinning thread (every few thousand polls): true = stop waiting (for good)
__device__ __forceinline__ bool rec_spin_check(const RecWatch& w, long long& t0, int kind, int step) {
    if (ld_relaxed_gpu(w.flag) != 0u) return true;
    const long long now = clock64();
    if (t0 == 0) { t0 = now; return false; }
    if (now - t0 <= w.spin_cycles) return false;
    rec_give_up(w.flag, w.host, kind, step);
    return true;
}

__device__ __forceinline__ void bounded_mbar_wait(uint64_t* bar, uint32_t parity, const RecWatch& w, bool& dead, int kind,
                                                  int step) {
    if (dead) return;
    uint32_t n = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++n & 0xFFFu) == 0 && rec_spin_check(w, t0, kind, step)) { dead = true; return; }
    }
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}


// Programmatic dependent launch: once EVERY CTA of this grid has executed this (i.e. the whole persistent grid is
// resident), a kernel enqueued behind it with the programmatic-serialization attribute may start on the SMs this grid
// leaves idle (148 - 125 / 128).  Kernels launched normally behind it are unaffected.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The other side: when THIS kernel was launched with the programmatic-serialization attribute behind a kernel that
// triggers early (the tensor-core GEMMs do), its CTAs become resident -- mbarriers, TMEM, the resident weight slice on
// its way -- while that kernel is still running; every thread that reads global memory the predecessor wrote calls this
// first.  Returns at once in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Profiling only (ZRB_REC_TRACE): the 8 launch slots in front of the per-step stamps.  Called by thread 0 of EVERY CTA at
// kernel entry (exit = false) and as its last instruction (exit = true):
//   [0]/[1] CTA 0's clock64 at entry / exit      [2]/[3] CTA 0's %globaltimer (ns) at entry / exit
//   [4] max over CTAs of -(entry %globaltimer)   [5] max over CTAs of the exit %globaltimer   (the host presets both
//   to the most negative value before each launch) -> [5] + [4] = lifetime of the whole grid in ns
__device__ __forceinline__ void rec_launch_stamps(long long* slots, bool cta0, bool exit) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    if (cta0) { slots[exit ? 1 : 0] = clock64(); slots[exit ? 3 : 2] = gt; }
    atomicMax(slots + (exit ? 5 : 4), exit ? gt : -gt);
}

// sigmoid / tanh on the SFU exp path (abs error ~1e-7, far below the fp16 operand noise of this engine)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// Publish this CTA's global writes of the step and arrive on the grid barrier: one release-RED at gpu scope
// (the bar.sync before it made the other epilogue threads' writes visible to this thread; release is
// cumulative).  The consumers' TMA reads are ordered by THEIR acquire + fence.proxy.async.
__device__ __forceinline__ void grid_counter_arrive(unsigned int* counter) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}

// spin on a global counter (grid barrier): relaxed polls (a plain L2 round trip each; ld.acquire would add an L1
// invalidate per poll), one acquire fence after the last arrival was seen; bounded like the mbarrier waits
__device__ __forceinline__ void grid_counter_wait(const unsigned int* counter, unsigned int target, const RecWatch& w,
                                                  bool& dead, int step) {
    if (dead) return;
    uint32_t n = 0;
    long long t0 = 0;
    while (ld_relaxed_gpu(counter) < target) {
        if ((++n & 0x3FFu) == 0 && rec_spin_check(w, t0, kWaitGrid, step)) { dead = true; return; }
    }
    asm volatile("fence.acquire.gpu;" ::: "memory");
}

}  // namespace zrb
